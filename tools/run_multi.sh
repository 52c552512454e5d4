# usage: bash tools/run_multi.sh N [extra bench args]   — torchrun launch of bench.py on N GPUs of this box
N=$1; shift
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_max_rel","n_gpus"): print(k, d.get(k))
    print("kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
    if "c4" in d: print("c4", json.dumps(d["c4"], indent=1)[:6000])
except Exception as e: print("no json:", e)
PY
