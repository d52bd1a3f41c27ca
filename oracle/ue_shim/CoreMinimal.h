/*
 * CoreMinimal.h -- stand-in for the part of Unreal Engine 5.4 that the reference's hot-path sources touch.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  It exists so that the reference's OWN code -- its headers included where they lie
 * under /root/reference, its function bodies extracted unmodified at build time by oracle/extract_ue_bodies.py -- can be
 * compiled without the engine (oracle/Makefile target `ref_ue` -> oracle/_ref/libref_ue_bodies.so) and used to pin
 * oracle/fs_oracle.c.  Nothing here is copied from the reference or from the engine: every type below is the smallest
 * std:: based object with the member names the reference uses.
 *
 * What stands in for the engine (the third-party dependency that is absent from /root/reference):
 *   UWorld::LineTraceSingleByObjectType   closest hit over a triangle list, double precision, brute force
 *   FMath::FRand / VRand / VRandCone      Philox4x32-10 keyed like the harness: one block per (work index, bounce, side);
 *                                         FRand = u0, VRand = uniform sphere(u1, u2), VRandCone(n, 90 deg) = cosine
 *                                         hemisphere(u1, u2) -- the FIX disposition of SURVEY 8a/A3 enters HERE, at the
 *                                         engine boundary, not in the reference's code
 * FVector is double precision as in UE 5.
 */
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <initializer_list>
#include <string>
#include <unordered_map>
#include <vector>

typedef int32_t int32;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uint8_t uint8;

#ifndef PI
#define PI (3.1415926535897932f)             /* UE: a float literal */
#endif
#define KINDA_SMALL_NUMBER (1.e-4f)
#define TEXT(x) x
#define UE_LOG(...) do { } while (0)
#define check(x) do { if (!(x)) ue_shim_check_failed(#x, __FILE__, __LINE__); } while (0)
#define UCLASS(...)
#define USTRUCT(...)
#define UPROPERTY(...)
#define UFUNCTION(...)
#define GENERATED_BODY()
#define GENERATED_USTRUCT_BODY()
#define FREQUENSEE_API
#define RETURN_QUICK_DECLARE_CYCLE_STAT(a, b) return TStatId()

void ue_shim_check_failed(const char* what, const char* file, int line);

template <typename T> T&& MoveTemp(T& v) { return static_cast<T&&>(v); }

typedef std::string FString;
struct FName { };
struct FColor { uint8 R, G, B, A; };
struct TStatId { };
struct FSubsystemCollectionBase { };
struct FAudioDevice;
struct FActorComponentTickFunction;
enum ELevelTick { LEVELTICK_All };
struct FPaths { static FString ProjectContentDir() { return FString(); } };

struct FMemory {
    static void* Memset(void* d, int v, size_t n) { return memset(d, v, n); }
    static void* Memcpy(void* d, const void* s, size_t n) { return memcpy(d, s, n); }
};

template <typename T>
class TArray {
public:
    TArray() {}
    TArray(std::initializer_list<T> il) : v(il) {}
    int32 Num() const { return (int32)v.size(); }
    bool IsEmpty() const { return v.empty(); }
    T& operator[](int32 i) { return v[(size_t)i]; }
    const T& operator[](int32 i) const { return v[(size_t)i]; }
    T* GetData() { return v.data(); }
    const T* GetData() const { return v.data(); }
    int32 Add(const T& x) { v.push_back(x); return (int32)v.size() - 1; }
    int32 AddUnique(const T& x) { for (size_t i = 0; i < v.size(); ++i) if (v[i] == x) return (int32)i; return Add(x); }
    int32 Remove(const T& x) { size_t n = v.size(); v.erase(std::remove(v.begin(), v.end(), x), v.end()); return (int32)(n - v.size()); }
    void AddDefaulted(int32 n) { v.resize(v.size() + (size_t)n); }
    void Append(const TArray<T>& o) { v.insert(v.end(), o.v.begin(), o.v.end()); }
    void Reserve(int32 n) { v.reserve((size_t)n); }
    void Reset(int32 n = 0) { v.clear(); v.reserve((size_t)n); }
    void Empty() { v.clear(); }
    void SetNum(int32 n) { v.resize((size_t)n); }
    /* engine semantics: existing elements are KEPT, only new ones are zero (so UFrequenSeeAudioComponent::FlushEnergyBuffer,
     * COMP.h:76-79, clears nothing once the buffer has its size -- see DESIGN.md section 2) */
    void SetNumZeroed(int32 n) { v.resize((size_t)n, T()); }
    void SetNumUninitialized(int32 n) { v.resize((size_t)n); }
    void Init(const T& x, int32 n) { v.assign((size_t)n, x); }
    T& Last() { return v.back(); }
    const T& Last() const { return v.back(); }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
    typename std::vector<T>::const_iterator begin() const { return v.begin(); }
    typename std::vector<T>::const_iterator end() const { return v.end(); }
    std::vector<T> v;
};

template <typename T>
class TWeakObjectPtr {
public:
    TWeakObjectPtr() : p(nullptr) {}
    TWeakObjectPtr(T* q) : p(q) {}
    TWeakObjectPtr(std::nullptr_t) : p(nullptr) {}
    bool IsValid() const { return p != nullptr; }
    T* Get() const { return p; }
    T* operator->() const { return p; }
    bool operator==(const TWeakObjectPtr& o) const { return p == o.p; }
private:
    T* p;
};

template <typename T>
class TObjectPtr {
public:
    TObjectPtr() : p(nullptr) {}
    TObjectPtr(T* q) : p(q) {}
    TObjectPtr(std::nullptr_t) : p(nullptr) {}
    operator bool() const { return p != nullptr; }
    T* operator->() const { return p; }
    T* Get() const { return p; }
private:
    T* p;
};

struct FVector {
    double X, Y, Z;
    FVector() : X(0), Y(0), Z(0) {}
    FVector(double x, double y, double z) : X(x), Y(y), Z(z) {}
    static const FVector ZeroVector;
    FVector operator+(const FVector& o) const { return FVector(X + o.X, Y + o.Y, Z + o.Z); }
    FVector operator-(const FVector& o) const { return FVector(X - o.X, Y - o.Y, Z - o.Z); }
    FVector operator*(double s) const { return FVector(X * s, Y * s, Z * s); }
    double Size() const { return sqrt(X * X + Y * Y + Z * Z); }
    FVector GetSafeNormal() const { const double l = Size(); return l > 1e-8 ? FVector(X / l, Y / l, Z / l) : FVector(); }
    bool IsNearlyZero() const { return fabs(X) <= 1e-4 && fabs(Y) <= 1e-4 && fabs(Z) <= 1e-4; }
    static double Dist(const FVector& a, const FVector& b) { return (a - b).Size(); }
    static double DotProduct(const FVector& a, const FVector& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
};
inline FVector operator*(double s, const FVector& v) { return v * s; }

/* ---- object model ---------------------------------------------------------------------------------------------- */
class UWorld;
class AActor;
class UObject { public: virtual ~UObject() {} };
class UDataAsset : public UObject { };
class UStaticMeshComponent : public UObject { };

class UActorComponent : public UObject {
public:
    virtual void OnRegister() {}
    virtual void OnUnregister() {}
    virtual void BeginPlay() {}
    virtual void TickComponent(float, ELevelTick, FActorComponentTickFunction*) {}
    AActor* GetOwner() const { return Owner; }
    UWorld* GetWorld() const { return World; }
    AActor* Owner = nullptr;
    UWorld* World = nullptr;
    struct { bool bCanEverTick; } PrimaryComponentTick = {false};
    bool bAutoActivate = false;
};
class UAudioComponent : public UActorComponent {
public:
    void FadeOut(float, float) {}
    bool bOverrideAttenuation = false;
    struct { bool bEnableOcclusion; } AttenuationOverrides = {false};
};

class UAcousticGeometryComponent;
class UFrequenSeeAudioComponent;

/* an actor = a location plus the two component kinds the path looks up */
class AActor : public UObject {
public:
    FVector Location;
    UAcousticGeometryComponent* Geometry = nullptr;
    UFrequenSeeAudioComponent* Audio = nullptr;
    FVector GetActorLocation() const;                 /* also the engine-side hook that starts a new random stream (see ue_shim_rng) */
    template <typename T> T* GetComponentByClass() const { return ue_shim_component((T*)nullptr); }
    template <typename T> T* FindComponentByClass() const { return ue_shim_component((T*)nullptr); }
private:
    UStaticMeshComponent* ue_shim_component(UStaticMeshComponent*) const { return nullptr; }
    UAcousticGeometryComponent* ue_shim_component(UAcousticGeometryComponent*) const { return Geometry; }
    UFrequenSeeAudioComponent* ue_shim_component(UFrequenSeeAudioComponent*) const { return Audio; }
};
class APawn : public AActor { };
class ADefaultPawn : public APawn { };

/* ---- scene queries ------------------------------------------------------------------------------------------------ */
enum ECollisionChannel { ECC_Pawn, ECC_WorldStatic, ECC_WorldDynamic };
struct FCollisionObjectQueryParams { void AddObjectTypesToQuery(ECollisionChannel) {} };
struct FCollisionQueryParams {
    FCollisionQueryParams() {}
    FCollisionQueryParams(const char*, bool) {}
    void AddIgnoredActor(const AActor*) {}
    void AddIgnoredComponent(const UStaticMeshComponent*) {}
};
struct FHitResult {
    FVector ImpactPoint, ImpactNormal;
    AActor* Actor = nullptr;
    AActor* GetActor() const { return Actor; }
};

/* the world: a triangle list (double precision), one actor per material */
class UWorld {
public:
    std::vector<double> v0, e1, e2, nrm;          /* [T][3] each */
    std::vector<uint32_t> tri_actor;              /* [T] index into Actors */
    std::vector<AActor*> Actors;
    uint64_t n_traces = 0;
    /* true when the segment Start -> End hits a triangle; closest hit, ties by lower triangle index */
    bool LineTraceSingleByObjectType(FHitResult& H, const FVector& Start, const FVector& End,
                                     const FCollisionObjectQueryParams&, const FCollisionQueryParams& = FCollisionQueryParams());
};

class UWorldSubsystem : public UObject {
public:
    virtual void Initialize(FSubsystemCollectionBase&) {}
    virtual void Deinitialize() {}
    UWorld* GetWorld() const { return World; }
    UWorld* World = nullptr;
};
class FTickableGameObject {
public:
    virtual ~FTickableGameObject() {}
    virtual void Tick(float) {}
    virtual TStatId GetStatId() const { return TStatId(); }
};

/* ---- FMath --------------------------------------------------------------------------------------------------------- */
struct FMath {
    static float Sqrt(float x) { return sqrtf(x); }
    static int32 CeilToInt(float x) { return (int32)ceilf(x); }
    static int32 FloorToInt(float x) { return (int32)floorf(x); }
    template <typename T> static T Clamp(T x, T lo, T hi) { return x < lo ? lo : (x < hi ? x : hi); }
    template <typename T> static T Min(T a, T b) { return a < b ? a : b; }
    static float Cos(float x) { return cosf(x); }
    static float Square(float x) { return x * x; }
    static float DegreesToRadians(float d) { return d * (PI / 180.f); }
    static uint32 RoundUpToPowerOfTwo(uint32 x) { uint32 p = 1; while (p < x) p <<= 1; return p; }
    static float FRand();
    static FVector VRand();
    static FVector VRandCone(const FVector& Dir, float ConeHalfAngleRad);
};

/* engine-side random streams: set by the driver, advanced by FRand (one Philox block per call) */
struct ue_shim_rng_state {
    uint64_t seed = 0;
    uint64_t index[2] = {0, 0};                  /* next work index per side (0: source actor, 1: listener pawn) */
    const AActor* side_actor[2] = {nullptr, nullptr};
    uint64_t g = 0; uint32_t side = 0, bounce = 0;
    uint32_t r[4] = {0, 0, 0, 0};
};
ue_shim_rng_state& ue_shim_rng();
