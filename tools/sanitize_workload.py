"""Small workload that touches every kernel family once (written for compute-sanitizer, which this GPU pool refuses to run --
profiles/r2l_sanitizer_unavailable.log; usable as `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_workload.py`
elsewhere, and run plainly by tests/test_gpu_workload.py): per-bounce pipeline,
persistent per-batch kernel, all-prefix connections, MIS, material model, shared listener, IR build (single / multi / per band),
convolver (single / multi), the multi-context reduce, the BVH build (PLOC) on a small furnished room."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-pathtracer_b200"))
import numpy as np
import frequensee as fs
from frequensee import scenes, capi

room = scenes.furnished_room(target_tris=12000)
box = scenes.shoebox()
rng = np.random.default_rng(0)
sums = []
for flags, env in ((0, {}), (0, {"FS_TUNE_MEGA": "1"}), (capi.FLAG_CONNECT_ALL, {}), (capi.FLAG_MIS, {}), (capi.FLAG_MATERIAL_MODEL, {}),
                   (capi.FLAG_SHARE_LISTENER, {}), (capi.FLAG_COUNT_VISITS, {})):
    os.environ.update(env)
    ctx = fs.Context(flags=flags)
    for k in env:
        os.environ.pop(k)
    M, B = room.absorption.shape
    if flags & capi.FLAG_MATERIAL_MODEL:
        ctx.set_scene_ex(room.verts, room.tri_mat, room.absorption, rng.uniform(0, .5, (M, B)).astype(np.float32),
                         rng.uniform(0, 1, (M, B)).astype(np.float32), rng.uniform(1, 9, M).astype(np.float32))
    else:
        ctx.set_scene(room.verts, room.tri_mat, room.absorption)
    src = np.array([room.sources[0], room.sources[0] + np.float32([.5, .3, .1])], np.float32)
    n = 300 if flags & (capi.FLAG_CONNECT_ALL | capi.FLAG_MIS) else 3000
    h = ctx.trace(src, room.listener, n, 8, 5)
    sums.append(int(h.sum() % (1 << 61)))
    ctx.build_ir(0); ctx.build_ir_all(2); ctx.build_ir_bands(7, 0, 1)
    ctx.conv_init_source(0); ctx.conv_init_source(1)
    x = rng.uniform(-.5, .5, (2, 1024, 2)).astype(np.float32)
    ctx.conv_process(x[0], 0); ctx.conv_process_multi(x, [0, 1]); ctx.conv_process_many(x, 1)
    ctx.close()
with fs.MultiContext([0, 0]) as m:
    m.set_scene(box.verts, box.tri_mat, box.absorption)
    sums.append(int(m.trace(box.sources, box.listener, 4000, 6, 3).sum() % (1 << 61)))
print("sanitize workload done", sums)
