#!/usr/bin/env python
"""Print the key figures of bench.py JSON lines: python tools/benchsum.py gpurun_out/x.json ..."""
import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:
        print(fn, "ERR", e)
        continue
    r = d.get("roofline") or {}
    e2e = (d.get("e2e") or {}).get("value")
    d.setdefault("fp32", {"frac_of_nominal": 0.0}); d.setdefault("clocks", {})
    print(f"{fn}: value={d['value']:.0f} ms/step={d['ms_per_step']:.3f} e2e={e2e and round(e2e)} "
          f"roof[{r.get('kernel')}]={r.get('frac', 0):.3f} share={r.get('share_of_step', 0):.2f} "
          f"fp32={d['fp32']['frac_of_nominal']:.3f} sm={d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
    print("    ", {k: round(v, 3) for k, v in (d.get("kernel_ms_per_step") or {}).items()})
