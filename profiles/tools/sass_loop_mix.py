"""Static instruction mix of the Riccati sweep loop of lqr_factor_dmma_warp_kernel<4,double> from the built object
(cuobjdump -sass): the loop is the largest backward branch whose body contains the DMMAs.  No GPU needed."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "chainer-differentiable-mpc_b200", "csrc", "build", "lqr_launch_f64.o")
fun = "_ZN4dmpc27lqr_factor_dmma_warp_kernelILi4EdEEvNS_9LqrParamsIT0_EE"
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), l))
loops = []
for a, op, l in ins:
    if op == "BRA":
        t = re.search(r"0x([0-9a-f]+)", l)
        if t and int(t.group(1), 16) < a:
            lo = int(t.group(1), 16)
            n_dmma = sum(1 for b, o, _ in ins if lo <= b <= a and o == "DMMA")
            loops.append((n_dmma, -(a - lo), lo, a))
n_dmma, _, lo, hi = max(loops)          # most DMMAs, then the tightest span that still holds them
c = collections.Counter(o for b, o, _ in ins if lo <= b <= hi)
print("sweep loop 0x%x..0x%x: %d instructions per element-step, %d DMMA" % (lo, hi, sum(c.values()), n_dmma))
for k, v in c.most_common(30):
    print("  %-10s %5d" % (k, v))
