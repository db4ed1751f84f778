"""Pretty-print the JSON line of a bench.py log (kernel table included)."""
import json
import sys

lines = [x for x in open(sys.argv[1]) if x.startswith("{")]
if not lines:
    print(open(sys.argv[1]).read()[-3000:])
    raise SystemExit(1)
d = json.loads(lines[-1])
ks = d.pop("kernels", [])
print(f"value {d['value'] / 1e6:.1f} Mpx/s  ms/step {d['ms_per_step']:.1f}  e2e {d['e2e']['value'] / 1e6 if d.get('e2e') else 0:.1f} Mpx/s  "
      f"launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
print("train", d.get("train_step"), "check", d.get("check"))
print("cpu", d.get("cpu_baseline"))
print("roofline", d.get("roofline"))
for k in ks:
    print(f"{k['kernel']:32s} ms={k['ms']:8.2f} n={k['launches']:4d} share={k['share_of_step']:.3f} tfl={k['tflops']:6.1f} "
          f"gbs={k['gbs']:7.0f} {k['bound']:6s} frac={k['frac']:.3f}")
