"""Print the headline numbers and the per-kernel-family profile of bench.py JSON lines: python tools/show_bench.py file..."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), d.get("fft_backend"), "row_mesh", d.get("row_mesh"),
              "launches", d.get("gpu_launches"))
        p = d.get("profile_one_realisation") or d.get("profile_one_step") or d.get("profile_one_pair")
        if p:
            tot = 0.0
            for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
                print("   %-36s %6d %9.3f" % (k, v["launches"], v["ms"]))
                tot += v["ms"]
            print("   total %.1f" % tot)
