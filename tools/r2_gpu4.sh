#!/bin/bash
# Round 2, GPU session 4: full parity suite incl. the reference-CUDA planes tests, burst A/B after the carve-out revert.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/r2_4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_4_pytest.log
tail -25 gpurun_out/r2_4_pytest.log
ab() {  # lib extra...
  lib=$1; shift
  if [ "$lib" = "default" ]; then unset RRT_B200_LIB; else export RRT_B200_LIB=$PWD/build/ab/librrt_$lib.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --depth 2 --share 1 "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $*', 'ms=%.2f'%d['ms_per_step'], 'alone_ms=%.2f'%d['latency_ms_single_frame'], 'steps/s=%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], d['clocks']['sm_mhz'])"
}
{
for lib in k0 k4 default k8two k16; do ab $lib; done
for lib in k0 default; do ab $lib --flags 0; done
for lib in k0 k4 default k8two k16; do ab $lib --camera C3; done
for lib in k0 default k8two; do ab $lib --camera C1; done
for lib in k0 default; do ab $lib --strict; done
for lib in k0 default; do ab $lib --width 1920 --height 1080; done
} 2>&1 | tee gpurun_out/r2_4_ab.log
unset RRT_B200_LIB
python - > gpurun_out/r2_4_density_err.txt 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, 'tests')
import relativisticraytracer_b200 as rrt
from oracle import Oracle
from inputs import disk_points
r = rrt.Renderer(0); ora = Oracle("port")
for flags in (3, 7):
    for t in (0.0, 1.0, 12.5):
        for name, seed in (("disk", 40), ("dust", 41)):
            q = disk_points(seed=seed)
            a = getattr(r, name + "_density")(rrt.default_params(flags=flags), q, t)
            b = getattr(ora, name + "_density")(ora.default_params(flags=flags), q, t)
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            print(name, "flags", flags, "t", t, "q99 %.3g q999 %.3g max %.3g zero-mismatch %d" % (np.quantile(err, .99), np.quantile(err, .999), err.max(), int(((a == 0) != (b == 0)).sum())))
PY
cat gpurun_out/r2_4_density_err.txt
