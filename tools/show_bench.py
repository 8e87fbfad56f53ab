"""Compact view of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j.get("roofline", {})
print(f"headline {j['metric']}: {j['value']/1e6:.2f} M tok/s, {j['ms_per_step']*1e3:.1f} us/step, n_gpus {j['n_gpus']}, launches {j.get('gpu_launches')}")
if r:
    print(f"  sweep kernel {r.get('kernel_ms',0)*1e3:.1f} us = {r.get('achieved',0):.0f} TF/s: frac burst {r.get('frac_burst',0):.3f} sustained {r.get('frac_sustained',0):.3f}; step frac burst {r['step']['frac_burst']:.3f} sustained {r['step']['frac_sustained']:.3f}")
e = j.get("e2e") or {}
if e:
    print(f"  e2e {e.get('value',0)/1e6:.2f} M (sync {e.get('synchronous',{}).get('value',0)/1e6:.2f} M, bf16 io {e.get('bf16_io',{}).get('value',0)/1e6:.2f} M)")
for k in ("exact_mode", "direct_launches", "cpu_baseline"):
    if k in j:
        print(f"  {k}: {j[k]['value']/1e6:.4f} M")
print("  clocks", j.get("clocks"), "parity", j.get("config", {}).get("parity_checked"))
def walk(name, c, ind=2):
    if not isinstance(c, dict):
        return
    if "value" in c and "ms_per_step" in c:
        rf = c.get("roofline", {})
        extra = ""
        if rf:
            extra = f" frac_b {rf.get('frac_burst', rf.get('frac', 0)):.3f} frac_s {rf.get('frac_sustained', 0):.3f}"
            if "step" in rf:
                extra += f" | step frac_s {rf['step']['frac_sustained']:.3f}"
        ex = c.get("exact_mode")
        print(" " * ind + f"{name}: {c['value']/1e6:.2f} M, {c['ms_per_step']*1e3:.1f} us{extra}" +
              (f", exact {ex['value']/1e6:.2f} M" if ex else "") + (f", e2e {c['e2e']['value']/1e6:.2f} M" if c.get('e2e') else "") +
              (f", parity {c['parity_checked']}" if 'parity_checked' in c else "") + (f", comm {c['comm_ms']*1e3:.0f} us" if 'comm_ms' in c else ""))
    elif "ms" in c and "roofline" in c:
        print(" " * ind + f"{name}: {c['ms']*1e3:.1f} us, {c['roofline']['achieved']:.0f} {c['roofline']['unit']} frac {c['roofline']['frac']:.2f}")
    elif "ms" in c and "tokens_per_s" in c:
        print(" " * ind + f"{name}: {c['ms']*1e3:.1f} us, {c['tokens_per_s']/1e6:.2f} M tokens/s")
    for k, v in c.items():
        if isinstance(v, dict) and k not in ("roofline", "exact_mode", "e2e", "clocks"):
            walk(k, v, ind + 2)
for name, c in (j.get("configs") or {}).items():
    walk(name, c)
for name, c in (j.get("variants") or {}).items():
    walk(name, c)
