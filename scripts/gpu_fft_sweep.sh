# sweep FFT-kernel launch shapes on config[1] and config[3] (env knobs in apb_plan_create)
mkdir -p gpurun_out
one() { # label, env...
  label=$1; shift
  for wl in c2 c4; do
    if [ $wl = c2 ]; then A="--steps 60 --warmup 5"; else A="--workload c4 --steps 10 --warmup 3"; fi
    env "$@" timeout 600 python bench.py $A --no-cpu > gpurun_out/sweep_${label}_$wl.json 2> gpurun_out/sweep_${label}_$wl.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_${label}_$wl.json").read().strip().splitlines()[-1]); km=d["kernel_ms"]; n=d["steps"]
    print("$label $wl", round(d["value"],2), "it/s", {k:round(km[k]["ms"]/n,3) for k in ("k_fft_cols","k_fft_rows","k_fft_rows_inv") if k in km})
except Exception as e:
    print("$label $wl FAILED", e, open("gpurun_out/sweep_${label}_$wl.err").read()[-300:])
PY
  done
}
one base APB_X=0
one c128 APB_FFT_NT_COLS=128 APB_FFT_COL_KB=55
one c128b APB_FFT_NT_COLS=128
one r128 APB_FFT_NT_ROWS=128 APB_FFT_NF_MAX=1
one r128b APB_FFT_NT_ROWS=128 APB_FFT_NF_MAX=2
