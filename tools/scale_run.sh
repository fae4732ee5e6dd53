#!/bin/bash
# bench.py on 1 / 2 / 4 / 8 GPUs of one box: default workload (--no-extras), configs[4] (ddqn), configs[2] (toroidal-regen).
# Usage (from the repo root, on an 8-GPU box): bash tools/scale_run.sh <tag>   -> gpurun_out/<tag>_{bench,ddqn,regen}_n{1,2,4,8}.json
tag=${1:-scale}
mkdir -p gpurun_out
run() {   # name, n, extra args...
    name=$1; n=$2; shift 2
    if [ "$n" = 1 ]; then
        timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/${tag}_${name}_n1.json 2> gpurun_out/${tag}_${name}_n1.err
    else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" \
            > gpurun_out/${tag}_${name}_n$n.json 2> gpurun_out/${tag}_${name}_n$n.err
    fi
    echo "$name n=$n rc=$? $(tail -c 300 gpurun_out/${tag}_${name}_n$n.json | head -c 0)$(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}_n$n.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(f"value {d['value']:.4g} ms/step {d['ms_per_step']:.4g} e2e {e.get('value')} allreduce_share {d.get('allreduce_share')}")
except Exception as ex:
    print("unreadable:", ex)
PY
)"
}
for n in 1 2 4 8; do run bench $n --steps 300 --warmup 50 --no-extras; done
for n in 1 2 4 8; do run ddqn $n --workload ddqn --steps 300 --warmup 30; done
for n in 1 8; do run regen $n --workload toroidal-regen --steps 300 --warmup 30; done
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
