set -e
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --batch 524288 --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-cli-simulator > gpurun_out/ovl_$name.json 2> gpurun_out/ovl_$name.err || { tail -3 gpurun_out/ovl_$name.err; }
  python - <<PY
import json
d=json.loads(open("gpurun_out/ovl_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"]/1e6,3), "M frames/s", d["ms_per_step"], d["counters"]["frames_ok"], d.get("roofline",{}).get("kernel_ms_per_launch"))
PY
}
run off RIA_OFDM_NO_OVERLAP=1
run d44 X=1
run c45 RIA_OFDM_OVERLAP_CTAS=4,5
run c35 RIA_OFDM_OVERLAP_CTAS=3,5
run c54 RIA_OFDM_OVERLAP_CTAS=5,4
run c36 RIA_OFDM_OVERLAP_CTAS=3,6
