#!/bin/bash
# Everything the round wants to see on a box with N GPUs (default 8): the reference's main.cpp on the GPUs, the P2P script,
# the C++ entry point tests, bench.py at N, the sweep for the plots, the host-path probe. Output under gpurun_out/<tag>_*.
N=${1:-8}; TAG=${2:-r2h}
python tools/write_cfg_mtx.py cfg2 /tmp/cfg2.mtx > gpurun_out/${TAG}_setup.log 2>&1
for p in 1 2 4 8; do
  [ $p -le $N ] || continue
  ./sparsematrixmultiplicationmpi_b200/bin/spmm_main -np $p 64 /tmp/cfg2.mtx > gpurun_out/${TAG}_main_np$p.log 2>&1; echo "spmm_main np=$p rc=$?"
  grep -E "Execution time|Results" gpurun_out/${TAG}_main_np$p.log | grep -v PETSc | tr "\n" ";"; echo
done
./oracle/_ref/ref_main -np $N 64 /tmp/cfg2.mtx > gpurun_out/${TAG}_refmain_np$N.log 2>&1
echo "reference main on the host cores, -np $N:"; grep -E "Execution time" gpurun_out/${TAG}_refmain_np$N.log | grep -v PETSc | tr "\n" ";"; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_p2p.py > gpurun_out/${TAG}_p2p.log 2>&1; tail -2 gpurun_out/${TAG}_p2p.log
python -m pytest tests/test_cxx_entry_points.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench$N.json"))
    print("bench N=$N:", d["value"], d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "first", d["e2e_first_call"]["ms"], "parity_ok", d.get("parity_ok"))
    print(json.dumps(d.get("north_star_scaling"))[:2500]); print(d.get("extras_error"))
except Exception as e:
    print("bench failed:", e)
PY
rm -f gpurun_out/${TAG}_results.csv
for p in 1 2 4 8; do
  [ $p -le $N ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $p --master-addr 127.0.0.1 --master-port 2953$p tools/sweep.py --matrices cfg2 --k 1,8,32,64 --out gpurun_out/${TAG}_results.csv --append > gpurun_out/${TAG}_sweep$p.log 2>&1; tail -1 gpurun_out/${TAG}_sweep$p.log
done
python tools/e2e_probe.py 64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(v,2) for k,v in d.items() if k.startswith('cxx')})"
