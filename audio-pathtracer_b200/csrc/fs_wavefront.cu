// fs_wavefront.cu -- wavefront BDPT: subpath extension, connection (shadow rays), path
// evaluation and fixed-point splatting.
//
// One batch = `batch` path pairs = 2*batch subpaths (side 0 from the source, side 1 from the
// listener; sp_id = 2*path + side).  Per bounce k one launch of k_extend consumes the queue of
// live subpaths and produces the next one, compacted with warp ballots (one atomic per warp), so
// warps stay full as Russian roulette / misses / max depth retire subpaths.  Kernels are
// persistent: grid = SMs x resident CTAs, each warp pulls 32 queue entries at a time from a
// device-side cursor, and queue lengths live on the device, so the whole update is enqueued
// without a host round trip.
//
//   k_extend  (GeneratePath, SUB.cpp:279-355)        RR, sample, closest hit, node record
//   k_connect (ConnectSubpaths, SUB.cpp:235-277)     any-hit ray between the two end nodes
//   k_eval    (EvaluatePath, SUB.cpp:360-420 + AddEnergyAtDelay, COMP.h:87-91)
//             energy product in path order, clamp/gain, Q32.32, warp-aggregated u64 atomics
//
// HBM layout (SoA, float4 granularity so every lane moves 16 B):
//   st_pos/st_nrm[2][2*cap]  ping-pong compacted subpath state
//   rec[(k)*2*cap + sp_id]   node record k>=1 of subpath sp_id: (segment length, material, pdf)
//   end_pos[sp_id]           last node position + node count
#include "fs_internal.h"

namespace {

constexpr int WF_THREADS = 256;

enum { MODE_BVH = 0, MODE_TOP = 1, MODE_BRUTE = 2 };

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// brute-force closest hit over the (sorted) triangle array; same (t, original id) rule
__device__ __forceinline__ int closest_brute(const fs_bvh_view& bv, fs_vec3 o, fs_vec3 d, float& best_t)
{
    int best = -1; uint32_t best_orig = 0xffffffffu;
    float bt = __int_as_float(0x7f800000);
    for (uint32_t i = 0; i < bv.n_tris; ++i) {
        const float4* tp = bv.tris + (size_t)i * 3;
        float4 a = fs_ldg4(tp), b = fs_ldg4(tp + 1), c = fs_ldg4(tp + 2);
        float t;
        if (fs_intersect_tri(o, d, fs_mk(a.x, a.y, a.z), fs_mk(b.x, b.y, b.z), fs_mk(c.x, c.y, c.z), t)) {
            uint32_t oi = __ldg(bv.tri_orig + i);
            if (t < bt || (t == bt && oi < best_orig)) { bt = t; best = (int)i; best_orig = oi; }
        }
    }
    best_t = bt;
    return best;
}
__device__ __forceinline__ bool any_brute(const fs_bvh_view& bv, fs_vec3 o, fs_vec3 d, float tmax)
{
    for (uint32_t i = 0; i < bv.n_tris; ++i) {
        const float4* tp = bv.tris + (size_t)i * 3;
        float4 a = fs_ldg4(tp), b = fs_ldg4(tp + 1), c = fs_ldg4(tp + 2);
        float t;
        if (fs_intersect_tri(o, d, fs_mk(a.x, a.y, a.z), fs_mk(b.x, b.y, b.z), fs_mk(c.x, c.y, c.z), t) && t < tmax)
            return true;
    }
    return false;
}

template <bool COUNT, int MODE>
__device__ __forceinline__ int trace_closest(const fs_trace_params& tp, const float4* top, fs_vec3 o, fs_vec3 d,
                                             float& t, fs_visit_counters* vc, uint32_t* overflow)
{
    if (MODE == MODE_BRUTE) return closest_brute(tp.bv, o, d, t);
    return fs_closest_hit<COUNT, MODE == MODE_TOP>(tp.bv, top, o, d, t, vc, overflow);
}
template <bool COUNT, int MODE>
__device__ __forceinline__ bool trace_any(const fs_trace_params& tp, const float4* top, fs_vec3 o, fs_vec3 d,
                                          float tmax, fs_visit_counters* vc, uint32_t* overflow)
{
    if (MODE == MODE_BRUTE) return any_brute(tp.bv, o, d, tmax);
    return fs_any_hit<COUNT, MODE == MODE_TOP>(tp.bv, top, o, d, tmax, vc, overflow);
}

__device__ __forceinline__ void stage_top(const fs_trace_params& tp, float4* smem_top)
{
    for (uint32_t i = threadIdx.x; i < tp.n_top * 4u; i += blockDim.x) smem_top[i] = tp.top[i];
    __syncthreads();
}

__device__ __forceinline__ void flush_counters(fs_dev_counters* dc, fs_visit_counters vc, bool shadow = false)
{
    uint32_t n = vc.nodes, t = vc.tris;
    for (int o = 16; o; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane_id() == 0) {
        atomicAdd(&dc->node_visits, (unsigned long long)n);
        atomicAdd(&dc->tri_tests, (unsigned long long)t);
        if (shadow) {
            atomicAdd(&dc->shadow_node_visits, (unsigned long long)n);
            atomicAdd(&dc->shadow_tri_tests, (unsigned long long)t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_extend: bounce k of every live subpath
// ---------------------------------------------------------------------------------------------
template <bool COUNT, int MODE>
__global__ void __launch_bounds__(WF_THREADS)
k_extend(const fs_trace_params tp, const fs_wave_buffers wb, uint32_t k, int in_buf,
         fs_dev_counters* __restrict__ dc)
{
    extern __shared__ float4 smem_top[];
    if (MODE == MODE_TOP) stage_top(tp, smem_top);
    const uint32_t lane = lane_id();
    const uint32_t count = (k == 0) ? 2u * tp.batch : wb.q_count[k];
    const float4* __restrict__ in_pos = wb.st_pos[in_buf];
    const float4* __restrict__ in_nrm = wb.st_nrm[in_buf];
    float4* __restrict__ out_pos = wb.st_pos[in_buf ^ 1];
    float4* __restrict__ out_nrm = wb.st_nrm[in_buf ^ 1];
    const uint32_t stride = 2u * wb.cap;
    fs_visit_counters vc; vc.nodes = 0; vc.tris = 0;
    uint32_t rays_local = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&wb.q_cursor[k], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const uint32_t j = base + lane;
        const bool valid = j < count;
        bool alive = false;
        fs_vec3 pos = fs_mk(0.f, 0.f, 0.f), nrm = fs_mk(0.f, 0.f, 0.f);
        uint32_t sp_id = 0;
        if (valid) {
            if (k == 0) {
                // node 0: actor location, zero normal, no material, probability 1 (SUB.cpp:287-291)
                sp_id = j;
                if (sp_id & 1u) {
                    pos = fs_mk(tp.lis[0], tp.lis[1], tp.lis[2]);
                } else {
                    uint64_t g = tp.g_first + (sp_id >> 1);
                    uint32_t s = (uint32_t)(g / tp.n_paths);
                    pos = fs_mk(__ldg(tp.src_pos + 3 * s), __ldg(tp.src_pos + 3 * s + 1), __ldg(tp.src_pos + 3 * s + 2));
                }
            } else {
                float4 a = in_pos[j], b = in_nrm[j];
                pos = fs_mk(a.x, a.y, a.z); sp_id = __float_as_uint(a.w);
                nrm = fs_mk(b.x, b.y, b.z);
            }
            const uint64_t g = tp.g_first + (sp_id >> 1);
            uint32_t r[4];
            fs_philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), k, sp_id & 1u, tp.seed_lo, tp.seed_hi, r);
            const float u0 = fs_u01(r[0]), u1 = fs_u01(r[1]), u2 = fs_u01(r[2]);
            bool terminated = true;
            uint32_t n_nodes = k + 1;                    // nodes pushed so far (SUB.cpp:297-298)
            if (u0 < tp.rr_prob) {                       // SUB.cpp:301-302
                fs_vec3 dir; float prob;
                if (k == 0) {                            // SUB.cpp:306-311
                    dir = fs_sample_sphere(u1, u2);
                    prob = FS_INV_4PI * tp.rr_prob;
                } else {                                 // SUB.cpp:312-318 (FIX: true cosine lobe)
                    float ct;
                    dir = fs_sample_cos_hemisphere(nrm, u1, u2, ct);
                    prob = (ct * FS_INV_PI) * tp.rr_prob;
                }
                ++rays_local;
                float t;
                uint32_t ovf_dummy = 0;
                int tri = trace_closest<COUNT, MODE>(tp, smem_top, pos, dir, t, &vc, &ovf_dummy);
                if (ovf_dummy) dc->overflow = 1u;
                if (tri >= 0) {                          // SUB.cpp:343-348
                    const float4* tq = tp.bv.tris + (size_t)tri * 3;
                    fs_vec3 fn = fs_mk(fs_ldg4(tq).w, fs_ldg4(tq + 1).w, fs_ldg4(tq + 2).w);
                    if (fs_dot(fn, dir) > 0.0f) { fn.x = -fn.x; fn.y = -fn.y; fn.z = -fn.z; }
                    fs_vec3 np;
                    np.x = fmaf(tp.eps_offset, fn.x, fmaf(t, dir.x, pos.x));
                    np.y = fmaf(tp.eps_offset, fn.y, fmaf(t, dir.y, pos.y));
                    np.z = fmaf(tp.eps_offset, fn.z, fmaf(t, dir.z, pos.z));
                    fs_vec3 dl = fs_sub(np, pos);
                    float seg = sqrtf(fs_dot(dl, dl));
                    uint32_t mat = __ldg(tp.bv.tri_mat + tri);
                    wb.rec[(size_t)(k + 1) * stride + sp_id] =
                        make_float4(seg, __uint_as_float(mat), prob, 0.f);
                    pos = np; nrm = fn;
                    n_nodes = k + 2;
                    terminated = (k + 1 >= tp.max_depth);        // PARAM: ray budget per subpath
                }
                // miss: FIX -> subpath ends at the current node
            }
            if (terminated) wb.end_pos[sp_id] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(n_nodes));
            alive = !terminated;
        }
        // ballot compaction into the next queue: one atomic per warp
        const uint32_t m = __ballot_sync(0xffffffffu, alive);
        if (m) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(&wb.q_count[k + 1], (uint32_t)__popc(m));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (alive) {
                const uint32_t o = slot + __popc(m & ((1u << lane) - 1u));
                out_pos[o] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(sp_id));
                out_nrm[o] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
            }
        }
    }
    for (int o = 16; o; o >>= 1) rays_local += __shfl_xor_sync(0xffffffffu, rays_local, o);
    if (lane == 0 && rays_local) atomicAdd(&dc->ext_rays, (unsigned long long)rays_local);
    if (COUNT) flush_counters(dc, vc);
}

// ---------------------------------------------------------------------------------------------
// k_connect: one visibility ray per path pair between the two LAST nodes
// ---------------------------------------------------------------------------------------------
template <bool COUNT, int MODE>
__global__ void __launch_bounds__(WF_THREADS)
k_connect(const fs_trace_params tp, const fs_wave_buffers wb, fs_dev_counters* __restrict__ dc,
          fs_path_dbg* __restrict__ dbg)
{
    extern __shared__ float4 smem_top[];
    if (MODE == MODE_TOP) stage_top(tp, smem_top);
    const uint32_t lane = lane_id();
    const uint32_t qi = tp.max_depth + 1;          // cursor / counter slot of this stage
    fs_visit_counters vc; vc.nodes = 0; vc.tris = 0;
    uint32_t rays_local = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&wb.q_cursor[qi], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= tp.batch) break;
        const uint32_t p = base + lane;
        const bool valid = p < tp.batch;
        bool connected = false;
        float len = 0.f;
        if (valid) {
            float4 fe = wb.end_pos[2u * p], be = wb.end_pos[2u * p + 1u];
            fs_vec3 F = fs_mk(fe.x, fe.y, fe.z), Bp = fs_mk(be.x, be.y, be.z);
            fs_vec3 dl = fs_sub(Bp, F);
            len = sqrtf(fs_dot(dl, dl));
            float tmax = len - tp.eps_connect;              // SUB.cpp:253
            bool occluded = false;
            if (tmax > 0.0f) {
                float inv = 1.0f / len;
                fs_vec3 dir = fs_mk(dl.x * inv, dl.y * inv, dl.z * inv);
                ++rays_local;
                uint32_t ovf_dummy = 0;
                occluded = trace_any<COUNT, MODE>(tp, smem_top, F, dir, tmax, &vc, &ovf_dummy);
                if (ovf_dummy) dc->overflow = 1u;
            }
            connected = !occluded;
            if (dbg) {
                fs_path_dbg* q = dbg + p;
                q->n_src_nodes = __float_as_uint(fe.w); q->n_lis_nodes = __float_as_uint(be.w);
                q->connected = connected ? 1u : 0u; q->bin = -1; q->delay_s = 0.f; q->total_dist = 0.f;
                for (int b = 0; b < FS_MAX_BANDS; ++b) q->energy[b] = 0.f;
                q->src_end[0] = fe.x; q->src_end[1] = fe.y; q->src_end[2] = fe.z;
                q->lis_end[0] = be.x; q->lis_end[1] = be.y; q->lis_end[2] = be.z;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, connected);
        if (m) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(&wb.q_count[qi], (uint32_t)__popc(m));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (connected) {
                const uint32_t o = slot + __popc(m & ((1u << lane) - 1u));
                wb.conn_queue[o] = p;
                wb.conn_len[o] = len;
            }
        }
    }
    for (int o = 16; o; o >>= 1) rays_local += __shfl_xor_sync(0xffffffffu, rays_local, o);
    if (lane == 0 && rays_local) atomicAdd(&dc->shadow_rays, (unsigned long long)rays_local);
    if (COUNT) flush_counters(dc, vc, true);
}

// ---------------------------------------------------------------------------------------------
// k_eval: EvaluatePath over F nodes ++ reverse(B nodes), then the splat
// ---------------------------------------------------------------------------------------------
// Sum of `v` over the lanes of `peers` (all lanes holding the same histogram key), result valid
// in the group's lowest lane.  Tree reduction over an arbitrary lane subset: in every round the
// even-ranked lanes absorb their next higher peer, odd-ranked lanes drop out.
__device__ __forceinline__ void reduce_peers8(uint32_t peers, unsigned long long v[FS_MAX_BANDS], int nb)
{
    const uint32_t group = peers;
    const uint32_t lane = lane_id();
    uint32_t rel = __popc(peers & ((1u << lane) - 1u));
    peers &= (0xfffffffeu << lane);                       // peers above me
    while (__any_sync(group, peers)) {
        const int next = __ffs(peers);                    // 1-based, 0 = none
        for (int b = 0; b < nb; ++b) {
            unsigned long long t = __shfl_sync(group, v[b], next ? next - 1 : (int)lane);
            if (next) v[b] += t;
        }
        const bool done = rel & 1u;
        if (done) peers = 0;
        peers &= __ballot_sync(group, !done);
        rel >>= 1;
    }
}

__global__ void __launch_bounds__(WF_THREADS)
k_eval(const fs_trace_params tp, const fs_wave_buffers wb, unsigned long long* __restrict__ hist,
       fs_dev_counters* __restrict__ dc, fs_path_dbg* __restrict__ dbg)
{
    const uint32_t lane = lane_id();
    const uint32_t qi = tp.max_depth + 1;
    const uint32_t count = wb.q_count[qi];
    const uint32_t stride = 2u * wb.cap;
    const uint32_t NBr = tp.ep.n_bands;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&dc->connected, (unsigned long long)count);
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < count;
         base += gridDim.x * blockDim.x) {
        const uint32_t j = base + lane;
        const bool valid = j < count;
        uint32_t key = 0xffffffffu;
        unsigned long long q[FS_MAX_BANDS];
#pragma unroll
        for (int b = 0; b < FS_MAX_BANDS; ++b) q[b] = 0ull;
        uint32_t s = 0, bin = 0;
        if (valid) {
            const uint32_t p = wb.conn_queue[j];
            const float len = wb.conn_len[j];
            const uint32_t sf = 2u * p, sb = 2u * p + 1u;
            const uint32_t nf = __float_as_uint(wb.end_pos[sf].w);
            const uint32_t nb = __float_as_uint(wb.end_pos[sb].w);
            float E[FS_MAX_BANDS];
#pragma unroll
            for (int b = 0; b < FS_MAX_BANDS; ++b) E[b] = 1.0f;
            float total = 0.0f;
            const float* row = nullptr;           // node 0 of the source subpath: no material
            float prob = 1.0f;
            for (uint32_t i = 1; i < nf; ++i) {   // segments F_{i-1} -> F_i
                float4 r = wb.rec[(size_t)i * stride + sf];
                fs_eval_segment<FS_MAX_BANDS>(tp.ep, row, prob, r.x, total, E);
                row = tp.refl_over_pi + (size_t)__float_as_uint(r.y) * NBr;
                prob = r.z;
            }
            fs_eval_segment<FS_MAX_BANDS>(tp.ep, row, prob, len, total, E);     // connection F_last -> B_last
            for (uint32_t i = nb - 1; i >= 1; --i) {                             // B_i -> B_{i-1}
                float4 r = wb.rec[(size_t)i * stride + sb];
                fs_eval_segment<FS_MAX_BANDS>(tp.ep, tp.refl_over_pi + (size_t)__float_as_uint(r.y) * NBr,
                                              r.z, r.x, total, E);
            }
            const float delay = total / tp.sound_speed;                           // SUB.cpp:419
            const float fb = floorf((delay * 1000.0f) / tp.bin_ms);               // COMP.h:89
            if (!(fb >= 0.0f)) bin = 0;
            else if (fb >= (float)(tp.n_bins - 1)) bin = tp.n_bins - 1;
            else bin = (uint32_t)fb;
            const uint64_t g = tp.g_first + p;
            s = (uint32_t)(g / tp.n_paths);
            key = s * tp.n_bins + bin;
#pragma unroll
            for (int b = 0; b < FS_MAX_BANDS; ++b) {
                if (b < (int)NBr) {
                    float e = E[b];
                    e = (e < tp.energy_clamp) ? e : tp.energy_clamp;              // SUB.cpp:410
                    e = e * tp.energy_gain;                                       // SUB.cpp:413
                    q[b] = (unsigned long long)(e * 4294967296.0f);               // Q32.32
                    if (dbg) dbg[p].energy[b] = e;
                }
            }
            if (dbg) { dbg[p].bin = (int32_t)bin; dbg[p].delay_s = delay; dbg[p].total_dist = total; }
        }
        // splat: lanes with the same (source, bin) combine first, one RED.64 per band per group
        const uint32_t active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            unsigned long long* h = hist + ((size_t)s * NBr) * tp.n_bins + bin;
            if (tp.flags & FS_FLAG_NO_SPLAT_AGG) {
                for (uint32_t b = 0; b < NBr; ++b) atomicAdd(h + (size_t)b * tp.n_bins, q[b]);
            } else {
                const uint32_t peers = __match_any_sync(active, key);
                if (peers != (1u << lane)) reduce_peers8(peers, q, (int)NBr);
                if ((uint32_t)(__ffs(peers) - 1) == lane)
                    for (uint32_t b = 0; b < NBr; ++b) atomicAdd(h + (size_t)b * tp.n_bins, q[b]);
            }
        }
    }
}

// max_depth == 0: no ray is ever extended, both subpaths consist of node 0 only
__global__ void k_init_ends(const fs_trace_params tp, const fs_wave_buffers wb)
{
    uint32_t sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp >= 2u * tp.batch) return;
    float x, y, z;
    if (sp & 1u) { x = tp.lis[0]; y = tp.lis[1]; z = tp.lis[2]; }
    else {
        uint64_t g = tp.g_first + (sp >> 1);
        uint32_t s = (uint32_t)(g / tp.n_paths);
        x = tp.src_pos[3 * s]; y = tp.src_pos[3 * s + 1]; z = tp.src_pos[3 * s + 2];
    }
    wb.end_pos[sp] = make_float4(x, y, z, __uint_as_float(1u));
}

__global__ void k_reset_queues(uint32_t* q_count, uint32_t* q_cursor, uint32_t n, fs_dev_counters* dc, int reset_dc)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { q_count[i] = 0u; q_cursor[i] = 0u; }
    if (reset_dc && i == 0) {
        dc->ext_rays = 0; dc->shadow_rays = 0; dc->connected = 0; dc->node_visits = 0; dc->tri_tests = 0;
        dc->shadow_node_visits = 0; dc->shadow_tri_tests = 0;
        dc->overflow = 0;
    }
}

// single rays (intersector parity tests)
template <int MODE>
__global__ void __launch_bounds__(WF_THREADS)
k_debug_rays(const fs_trace_params tp, const float* __restrict__ rays, const float* __restrict__ tmax,
             uint64_t n, float* __restrict__ out_t, uint32_t* __restrict__ out_tri, uint8_t* __restrict__ out_hit,
             fs_dev_counters* __restrict__ dc)
{
    extern __shared__ float4 smem_top[];
    if (MODE == MODE_TOP) stage_top(tp, smem_top);
    fs_visit_counters vc; vc.nodes = 0; vc.tris = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        fs_vec3 o = fs_mk(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]);
        fs_vec3 d = fs_mk(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
        uint32_t ovf = 0;
        if (out_hit) {
            out_hit[i] = trace_any<false, MODE>(tp, smem_top, o, d, tmax[i], &vc, &ovf) ? 1 : 0;
        } else {
            float t;
            int tri = trace_closest<false, MODE>(tp, smem_top, o, d, t, &vc, &ovf);
            out_t[i] = t;
            out_tri[i] = tri >= 0 ? __ldg(tp.bv.tri_orig + tri) : 0xffffffffu;
        }
        if (ovf) dc->overflow = 1u;
    }
}

template <typename K>
int resident_ctas(K kernel, int threads, size_t smem)
{
    int n = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int pick_mode(const fs_trace_params& tp)
{
    if (tp.flags & FS_FLAG_BRUTE_FORCE) return MODE_BRUTE;
    if ((tp.flags & FS_FLAG_NO_TREELET) || tp.n_top < 2) return MODE_BVH;
    return MODE_TOP;
}

}  // namespace

cudaError_t fs_wave_alloc(fs_ctx* ctx, uint32_t cap, uint32_t max_depth)
{
    fs_wave_buffers* wb = &ctx->wb;
    if (wb->cap >= cap && wb->depth_cap >= max_depth && wb->rec) return cudaSuccess;
    if (cap < wb->cap) cap = wb->cap;
    if (max_depth < wb->depth_cap) max_depth = wb->depth_cap;
    fs_wave_free(wb);
    cudaError_t e;
    const size_t n2 = 2ull * cap;
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc(&wb->st_pos[i], sizeof(float4) * n2)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&wb->st_nrm[i], sizeof(float4) * n2)) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc(&wb->rec, sizeof(float4) * n2 * (max_depth + 1ull))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&wb->end_pos, sizeof(float4) * n2)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&wb->conn_queue, 4ull * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&wb->conn_len, 4ull * cap)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&wb->q_count, 4ull * (max_depth + 2))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&wb->q_cursor, 4ull * (max_depth + 2))) != cudaSuccess) return e;
    wb->cap = cap; wb->depth_cap = max_depth;
    return cudaSuccess;
}

void fs_wave_free(fs_wave_buffers* wb)
{
    for (int i = 0; i < 2; ++i) { cudaFree(wb->st_pos[i]); cudaFree(wb->st_nrm[i]); }
    cudaFree(wb->rec); cudaFree(wb->end_pos); cudaFree(wb->conn_queue); cudaFree(wb->conn_len);
    cudaFree(wb->q_count); cudaFree(wb->q_cursor);
    memset(wb, 0, sizeof(*wb));
}

template <bool COUNT, int MODE>
static cudaError_t launch_batch(fs_ctx* ctx, const fs_trace_params& tp, unsigned long long* d_hist,
                                fs_path_dbg* d_dbg)
{
    cudaStream_t st = ctx->stream;
    const fs_wave_buffers& wb = ctx->wb;
    const size_t smem = (MODE == MODE_TOP) ? (size_t)tp.n_top * 64 : 0;
    static int occ_ext = 0, occ_con = 0;
    if (!occ_ext) occ_ext = resident_ctas(k_extend<COUNT, MODE>, WF_THREADS, smem);
    if (!occ_con) occ_con = resident_ctas(k_connect<COUNT, MODE>, WF_THREADS, smem);
    if (smem > 48 * 1024) {
        cudaFuncSetAttribute(k_extend<COUNT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_connect<COUNT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const bool timing = (tp.flags & FS_FLAG_TIME_KERNELS) != 0;
    cudaEvent_t* ev = nullptr;
    if (timing) {
        if (ctx->kev.size() < ctx->kev_used + 4) {
            size_t old = ctx->kev.size();
            ctx->kev.resize(ctx->kev_used + 4);
            for (size_t i = old; i < ctx->kev.size(); ++i) cudaEventCreate(&ctx->kev[i]);
        }
        ev = &ctx->kev[ctx->kev_used];
        ctx->kev_used += 4;
    }
    const uint32_t nq = tp.max_depth + 2;
    k_reset_queues<<<(nq + 63) / 64, 64, 0, st>>>(wb.q_count, wb.q_cursor, nq, ctx->d_counters, 0);
    ++ctx->stats.kernel_launches;
    // persistent grids: SMs x resident CTAs, capped by the work available
    const uint32_t warps_needed = (2u * tp.batch + 31u) / 32u;
    uint32_t grid_ext = (uint32_t)(ctx->sm_count * occ_ext);
    uint32_t ctas_needed = (warps_needed + WF_THREADS / 32 - 1) / (WF_THREADS / 32);
    if (grid_ext > ctas_needed) grid_ext = ctas_needed ? ctas_needed : 1;
    if (timing) cudaEventRecord(ev[0], st);
    if (tp.max_depth == 0) {
        k_init_ends<<<(2u * tp.batch + 255u) / 256u, 256, 0, st>>>(tp, wb);
        ++ctx->stats.kernel_launches;
    }
    for (uint32_t k = 0; k < tp.max_depth; ++k) {
        k_extend<COUNT, MODE><<<grid_ext, WF_THREADS, smem, st>>>(tp, wb, k, (int)(k & 1u), ctx->d_counters);
        ++ctx->stats.kernel_launches;
        ++ctx->stats.extend_launches;
    }
    if (timing) cudaEventRecord(ev[1], st);
    uint32_t grid_con = (uint32_t)(ctx->sm_count * occ_con);
    uint32_t ctas_con = ((tp.batch + 31u) / 32u + WF_THREADS / 32 - 1) / (WF_THREADS / 32);
    if (grid_con > ctas_con) grid_con = ctas_con ? ctas_con : 1;
    k_connect<COUNT, MODE><<<grid_con, WF_THREADS, smem, st>>>(tp, wb, ctx->d_counters, d_dbg);
    ++ctx->stats.kernel_launches;
    if (timing) cudaEventRecord(ev[2], st);
    uint32_t grid_ev = (uint32_t)ctx->sm_count * 4u;
    if (grid_ev > ctas_con) grid_ev = ctas_con ? ctas_con : 1;
    k_eval<<<grid_ev, WF_THREADS, 0, st>>>(tp, wb, d_hist, ctx->d_counters, d_dbg);
    ++ctx->stats.kernel_launches;
    if (timing) cudaEventRecord(ev[3], st);
    return cudaGetLastError();
}

cudaError_t fs_wave_trace_batch(fs_ctx* ctx, const fs_trace_params& tp, unsigned long long* d_hist,
                                fs_path_dbg* d_dbg)
{
    const bool count = (tp.flags & FS_FLAG_COUNT_VISITS) != 0;
    switch (pick_mode(tp)) {
    case MODE_BRUTE: return launch_batch<false, MODE_BRUTE>(ctx, tp, d_hist, d_dbg);
    case MODE_TOP:
        return count ? launch_batch<true, MODE_TOP>(ctx, tp, d_hist, d_dbg)
                     : launch_batch<false, MODE_TOP>(ctx, tp, d_hist, d_dbg);
    default:
        return count ? launch_batch<true, MODE_BVH>(ctx, tp, d_hist, d_dbg)
                     : launch_batch<false, MODE_BVH>(ctx, tp, d_hist, d_dbg);
    }
}

// reset of the per-call device counters (first batch of a trace call)
cudaError_t fs_wave_reset_counters(fs_ctx* ctx)
{
    k_reset_queues<<<1, 64, 0, ctx->stream>>>(ctx->wb.q_count, ctx->wb.q_cursor, 0, ctx->d_counters, 1);
    ++ctx->stats.kernel_launches;
    return cudaGetLastError();
}

cudaError_t fs_wave_debug_rays(fs_ctx* ctx, const fs_trace_params& tp, const float* d_rays, const float* d_tmax,
                               uint64_t n, float* d_t, uint32_t* d_tri, uint8_t* d_hit)
{
    cudaStream_t st = ctx->stream;
    const int mode = pick_mode(tp);
    const size_t smem = (mode == MODE_TOP) ? (size_t)tp.n_top * 64 : 0;
    uint32_t grid = (uint32_t)((n + WF_THREADS - 1) / WF_THREADS);
    if (grid > (uint32_t)ctx->sm_count * 8u) grid = (uint32_t)ctx->sm_count * 8u;
    if (!grid) grid = 1;
    if (mode == MODE_BRUTE) k_debug_rays<MODE_BRUTE><<<grid, WF_THREADS, 0, st>>>(tp, d_rays, d_tmax, n, d_t, d_tri, d_hit, ctx->d_counters);
    else if (mode == MODE_TOP) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_debug_rays<MODE_TOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_debug_rays<MODE_TOP><<<grid, WF_THREADS, smem, st>>>(tp, d_rays, d_tmax, n, d_t, d_tri, d_hit, ctx->d_counters);
    } else k_debug_rays<MODE_BVH><<<grid, WF_THREADS, 0, st>>>(tp, d_rays, d_tmax, n, d_t, d_tri, d_hit, ctx->d_counters);
    ++ctx->stats.kernel_launches;
    return cudaGetLastError();
}
