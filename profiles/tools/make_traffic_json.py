#!/usr/bin/env python
"""profiles/r2/traffic.json from an `ncu --set full` capture of the three config-5 kernels.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv | python profiles/tools/make_traffic_json.py CAPTURE_NAME BATCH > profiles/r2/traffic.json

Stores dram__bytes_read.sum + dram__bytes_write.sum per launch, divided by the batch of the captured launch, and a
fingerprint of the kernel sources (comments and whitespace stripped) so that bench.py can tell when the kernels have changed
since the capture (`roofline.traffic_stale`)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
capture, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
K, R, W = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture %s (summary: "
                   "profiles/r2/%s.summary.txt; shipped kernels, B=%d, FACTOR|ROLLOUT|SAVE_FAC with V_t|v_t emission; two-sweep "
                   "adjoint), divided by the batch of the captured launch -> bytes per solve (config 5: n=32 m=8 T=100 fp64). "
                   "bench.py multiplies by the batch of the timed launch and compares _kernel_fingerprint with the sources it "
                   "runs." % (capture, capture, batch),
       "_kernel_fingerprint": bench.kernel_fingerprint()}
for r in rows[2:]:
    name = r[K]
    key = ("lqr_factor_dmma_warp_kernel" if "lqr_factor_dmma_warp_kernel" in name else
           "lqr_dtau_kernel_fused1" if "lqr_dtau_kernel" in name else
           "adjoint_fused_kernel" if "adjoint_fused_kernel" in name else None)
    if key is None or key in out:
        continue
    rd = float(r[R].replace(",", "")) * UNIT[units[R]]
    wr = float(r[W].replace(",", "")) * UNIT[units[W]]
    out[key] = {"bytes_per_solve": int(round((rd + wr) / batch)), "capture": capture, "read": int(rd), "write": int(wr), "batch": batch}
print(json.dumps(out, indent=1))
