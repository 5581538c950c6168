"""Prints the headline fields of bench.py JSON lines read from stdin."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    mc = d.get("equity_mc") or {}
    hb = (d.get("eval7_hbm") or {}).get("roofline") or {}
    cb = d.get("cpu_baseline") or {}
    print(f"impl={d.get('impl', 'b200')} N={d['n_gpus']} value={d['value']:.4g} {d['unit']} ms/step={d['ms_per_step']:.4f} "
          f"kernel_ms={r.get('kernel_ms')} frac={r.get('frac')} e2e={(d.get('e2e') or {}).get('value')} cpu={cb.get('value')} "
          f"mc={mc.get('value')} mc_ms={mc.get('ms_per_step')} mc_sum={mc.get('checksum')} hbm_frac={hb.get('frac')} "
          f"launches={d.get('gpu_launches')} clocks={d.get('clocks')}")
