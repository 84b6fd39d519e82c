#!/usr/bin/env python3
"""Extracts the known-answer vectors of the reference's own unit tests into tests/golden/kat.json
(run once in the build container, where /root/reference exists; the JSON is committed):
  * natural coefficient orders 8x8 / 16x8 (jxl/src/frame/coeff_order.rs:155-169, "golden from libjxl")
  * sampled library dequant matrices       (jxl/src/frame/quant_weights.rs:1232-2122, tolerance 1e-5)
Only numeric literals are read."""
import json, re
out = {}
src = open('/root/reference/jxl/src/frame/coeff_order.rs').read()
for name in ['COEFF_ORDER_1X1', 'COEFF_ORDER_2X1']:
    m = re.search(r'const %s: \[u32; \d+\] = \[(.*?)\];' % name, src, re.S)
    out[name] = [int(v) for v in m.group(1).replace('\n', ' ').split(',') if v.strip()]
q = open('/root/reference/jxl/src/frame/quant_weights.rs').read()
m = re.search(r'let target_table = \[(.*?)\];', q, re.S)
out['DEQUANT_TARGET_TABLE'] = [float(v.strip().replace('f32', '')) for v in m.group(1).split(',') if v.strip()]
json.dump(out, open('/root/repo/tests/golden/kat.json', 'w'))
print({k: len(v) for k, v in out.items()})
