#!/bin/bash
# compute-sanitizer on small cases of every K1 kernel (three-pass with bulk prefetch FP64 / FP32, general), K2 / K3, K5 / K6 and the staging
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from transport_analysis_b200.synthetic import make_universe, random_trajectory
from transport_analysis_b200.velocityautocorr import VelocityAutocorr
from transport_analysis_b200.viscosity import ViscosityHelfand
T, N, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
prec = "fp32" if mode.endswith("32") else "fp64"
mode = mode[:-2] if mode.endswith("32") else mode
vel, pos = random_trajectory(T, N, seed=1, with_positions=True, rho=0.5)
u = make_universe(pos, vel, masses=np.full(N, 12.0), dimensions=[20, 20, 20, 90, 90, 90])
if mode == "fft":
    a = VelocityAutocorr(u.atoms, fft=True, precision=prec).run(); print(a._ctx.fft_plan_info())
elif mode == "win":
    VelocityAutocorr(u.atoms, fft=False, precision=prec).run()
else:
    ViscosityHelfand(u.atoms, fft=(mode == "helfft")).run()
print("done", mode, T, N)
PY
for tool in memcheck racecheck synccheck; do
  for spec in "3000 300 fft" "10000 160 fft" "12000 150 fft" "5000 150 fft:general" "10000 160 fft32" "3000 300 fft32" "700 40 fft" "700 40 fft32" "600 60 win32" "600 60 win" "600 60 hel" "3000 20 helfft"; do
    set -- $spec; mode=${3%%:*}; path=""; [[ "$3" == *:* ]] && path=${3##*:}
    out=$(TA_B200_K1_PATH=$path compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py $1 $2 $mode 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" | head -4 | tr '\n' ' ')
    echo "$tool [$spec] $out"
  done
done
