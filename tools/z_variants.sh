#!/bin/bash
# A/B of the fused z pass variants (SWW_Z16_VARIANT) on the C2 workload: prints step time and the fused-Z time
for v in "$@"; do
  SWW_Z16_VARIANT=$v python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/var_$v.log 2>&1
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
for l in open("gpurun_out/var_%s.log"%v):
    if l.startswith("{"):
        d=json.loads(l); p=d["profile_one_realisation"]
        print("variant", v, "ms/step %.1f"%d["ms_per_step"], "Zfused %.1f"%p["fft.Z c2r*w*r2c fused"]["ms"], {k:v_["ms"] for k,v_ in p.items() if v_["ms"]>5})
P
done
