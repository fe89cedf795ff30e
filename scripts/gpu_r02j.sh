mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02j_c4_$tag.json 2> gpurun_out/r02j_c4_$tag.err; env "$@" timeout 600 python bench.py --workload c2 --steps 40 --warmup 3 --no-cpu > gpurun_out/r02j_c2_$tag.json 2> gpurun_out/r02j_c2_$tag.err; }
run r16 APB_X=1
run r8 APB_FFT_MAXR=8
run r4 APB_FFT_MAXR=4
run r4no9 APB_FFT_MAXR=4 APB_FFT_NO9=1
run r8t512 APB_FFT_MAXR=8 APB_FFT_NT_COLS=512 APB_FFT_NT_ROWS=512
run r16t128 APB_FFT_NT_COLS=128 APB_FFT_NT_ROWS=128
python - <<'PY'
import json
for wl in ("c4","c2"):
  for n in ("r16","r8","r4","r4no9","r8t512","r16t128"):
    try:
        d=json.loads(open(f"gpurun_out/r02j_{wl}_{n}.json").read().strip().splitlines()[-1])
        km=d["kernel_ms"]
        print(wl, n, round(d["ms_per_step"],3), {k:round(km[k]["ms"]/d["steps"],3) for k in ("k_fft_cols","k_fft_rows","k_fft_rows_inv")})
    except Exception as e:
        print(wl, n, "FAILED", e); print(open(f"gpurun_out/r02j_{wl}_{n}.err").read()[-600:])
PY
