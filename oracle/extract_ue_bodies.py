#!/usr/bin/env python
"""extract_ue_bodies.py -- build-time extraction of the reference's hot-path function bodies.

TEST INFRASTRUCTURE ONLY (oracle/).  Run by `make -C oracle ref_ue` when /root/reference is present.  It reads the
reference sources WHERE THEY LIE and writes, into a scratch directory that is deleted after the compile (nothing of
the reference is committed to this repo):

  sub_bodies.inc    UAudioRayTracingSubsystem::{UpdateSource, GenerateFullPaths, ConnectSubpaths, GeneratePath,
                    EvaluatePath}                                   Private/AudioRayTracingSubsystem.cpp
  comp_bodies.inc   UFrequenSeeAudioComponent::{ctor, dtor, ReconstructImpulseResponse, NormalizeImpulseResponse}
                                                                    Private/FrequenSeeAudioComponent.cpp
  rev_members.inc   the state members of FFrequenSeeAudioReverbPlugin  Private/FrequenSeeAudioReverbPlugin.h
  rev_bodies.inc    FFrequenSeeAudioReverbPlugin::{Initialize, ConvolveFFT}  Private/FrequenSeeAudioReverbPlugin.cpp
  <engine header names>   one-line forwarding headers to oracle/ue_shim/CoreMinimal.h, so that the reference's own
                    headers (AudioRayTracingSubsystem.h, FrequenSeeAudioComponent.h, AcousticMaterial.h,
                    AcousticGeometryComponent.h, CircularBuffer.h) can be #included in place

Every function is found by its qualified name and cut by brace matching -- UNMODIFIED text, verbatim; the script
fails loudly if a function is missing.  Each piece is preceded by a #line directive naming the reference file, so
compiler diagnostics and debuggers point at the reference, and a manifest lists file:first-last line of every piece.
"""
import os
import re
import sys

REF = "/root/reference/Plugins/FrequenSee/Source/FrequenSee"
SUB_CPP = REF + "/Private/AudioRayTracingSubsystem.cpp"
COMP_CPP = REF + "/Private/FrequenSeeAudioComponent.cpp"
REV_CPP = REF + "/Private/FrequenSeeAudioReverbPlugin.cpp"
REV_H = REF + "/Private/FrequenSeeAudioReverbPlugin.h"

STUB_HEADERS = [
    "Subsystems/WorldSubsystem.h", "GameFramework/DefaultPawn.h", "Components/ActorComponent.h",
    "Components/AudioComponent.h", "Engine/DataAsset.h", "Audio.h",
    "AudioRayTracingSubsystem.generated.h", "FrequenSeeAudioComponent.generated.h", "AcousticMaterial.generated.h",
    "AcousticGeometryComponent.generated.h",
]


def read(path):
    with open(path, encoding="utf-8-sig") as f:
        return f.read()


def strip_comments_keep_layout(text):
    """same length as text, comments and string/char literals blanked (so braces inside them are not counted)"""
    out = list(text)
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if text.startswith("//", i):
            j = text.find("\n", i)
            j = n if j < 0 else j
            for k in range(i, j):
                out[k] = " "
            i = j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            j = n if j < 0 else j + 2
            for k in range(i, j):
                if out[k] != "\n":
                    out[k] = " "
            i = j
        elif c in "\"'":
            j = i + 1
            while j < n and text[j] != c:
                j += 2 if text[j] == "\\" else 1
            for k in range(i + 1, min(j, n)):
                out[k] = " "
            i = j + 1
        else:
            i += 1
    return "".join(out)


def function_text(path, text, qualified):
    """verbatim text of the definition `... qualified(...) ... { ... }`, with its 1-based first/last line"""
    clean = strip_comments_keep_layout(text)
    m = None
    for cand in re.finditer(re.escape(qualified) + r"\s*\(", clean):
        # a definition: the matching ')' is followed (optionally by const / an initialiser list) by '{'
        depth, j = 0, cand.end() - 1
        while j < len(clean):
            if clean[j] == "(":
                depth += 1
            elif clean[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        k = j + 1
        tail = re.match(r"\s*(const)?\s*(:[^{;]*)?\{", clean[k:])
        if tail:
            m = (cand.start(), k + tail.end() - 1)
            break
    if m is None:
        raise SystemExit("extract_ue_bodies: definition of %s not found in %s" % (qualified, path))
    start_name, brace = m
    start = clean.rfind("\n", 0, start_name) + 1            # the whole first line (return type)
    depth, j = 0, brace
    while j < len(clean):
        if clean[j] == "{":
            depth += 1
        elif clean[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    if depth != 0:
        raise SystemExit("extract_ue_bodies: unbalanced braces in %s" % qualified)
    first = text.count("\n", 0, start) + 1
    last = text.count("\n", 0, j) + 1
    return text[start:j + 1], first, last


def region_text(path, text, first_pat, last_pat):
    a = re.search(first_pat, text)
    b = re.search(last_pat, text[a.end():]) if a else None
    if not a or not b:
        raise SystemExit("extract_ue_bodies: region %r .. %r not found in %s" % (first_pat, last_pat, path))
    start = text.rfind("\n", 0, a.start()) + 1
    end = a.end() + b.end()
    return text[start:end], text.count("\n", 0, start) + 1, text.count("\n", 0, end) + 1


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    manifest = []

    def emit(name, path, pieces):
        with open(os.path.join(out, name), "w") as f:
            for body, first, last in pieces:
                f.write('#line %d "%s"\n' % (first, path))
                f.write(body + "\n")
                manifest.append("%s <- %s:%d-%d" % (name, path, first, last))

    sub = read(SUB_CPP)
    emit("sub_bodies.inc", SUB_CPP, [function_text(SUB_CPP, sub, "UAudioRayTracingSubsystem::" + fn) for fn in
                                      ("UpdateSource", "GenerateFullPaths", "ConnectSubpaths", "GeneratePath", "EvaluatePath")])
    comp = read(COMP_CPP)
    emit("comp_bodies.inc", COMP_CPP, [function_text(COMP_CPP, comp, "UFrequenSeeAudioComponent::" + fn) for fn in
                                        ("UFrequenSeeAudioComponent", "~UFrequenSeeAudioComponent",
                                         "ReconstructImpulseResponse", "NormalizeImpulseResponse")])
    rev = read(REV_CPP)
    emit("rev_bodies.inc", REV_CPP, [function_text(REV_CPP, rev, "FFrequenSeeAudioReverbPlugin::" + fn) for fn in
                                      ("Initialize", "ConvolveFFT")])
    revh = read(REV_H)
    emit("rev_members.inc", REV_H, [region_text(REV_H, revh, r"int SamplingRate = 0;", r"void ConvolveFFT\([^;]*;")])
    for h in STUB_HEADERS:
        p = os.path.join(out, h)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write('#include "CoreMinimal.h"\n')
    with open(os.path.join(out, "MANIFEST.txt"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    print("\n".join(manifest))


if __name__ == "__main__":
    main()
