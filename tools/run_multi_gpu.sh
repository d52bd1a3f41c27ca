#!/bin/bash
# Multi-GPU measurements of one round (run under `gpurun --gpus 8`): the driver's weak-scaling line at N = 8 (with the
# north-star point, the NCCL parity check and the C-ABI multi-GPU path in the same JSON line), the strong-scaling curve
# (2^20 pairs in total over 1/2/4/8 GPUs), and the plain C++ multi-GPU smoke test.  usage: bash tools/run_multi_gpu.sh <tag>
T=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29520 + N)) bench.py --gpus $N --steps 20 --warmup 5 --strong --north-star off --no-cpu-baseline > gpurun_out/${T}_strong_n${N}.json 2> gpurun_out/${T}_strong_n${N}.err
done
python bench.py --gpus 1 --steps 20 --warmup 5 --strong --no-cpu-baseline > gpurun_out/${T}_strong_n1.json 2> gpurun_out/${T}_strong_n1.err
g++ -std=c++17 -O1 -I include tests/cpp/multi_smoke.cpp -o tests/cpp/multi_smoke -L audio-pathtracer_b200/lib -lfrequensee -Wl,-rpath,$PWD/audio-pathtracer_b200/lib
for N in 2 8; do ./tests/cpp/multi_smoke $N > gpurun_out/${T}_multi_smoke_n${N}.json 2>&1; done
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --scene mine_tunnels --sources 64 --paths 1048576 --depth 16 --north-star off > gpurun_out/${T}_bench_config4_tunnels_64sources_n8.json 2> gpurun_out/${T}_config4.err
python -m pytest tests/test_cpp_multi.py -m gpu -q 2>&1 | tail -2 > gpurun_out/${T}_pytest_multi.log
for f in gpurun_out/${T}_*.json; do echo "== $f"; head -c 600 $f; echo; done
tail -3 gpurun_out/${T}_bench_n8.err
