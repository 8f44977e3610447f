"""Where would time go?  Per-node roofline estimate of a lowered plan."""
import sys, math, time
sys.path.insert(0, ".")
from collections import defaultdict
from qibotn_b200 import circuits, executor
from qibotn_b200.network import QiboCircuitToEinsum
from qibotn_b200.observables import get_ham_gates
from qibotn_b200.planner import find_path

which = sys.argv[1]
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if which == "c2":
    c = circuits.brickwork(30, 12, seed=0)
    conv = QiboCircuitToEinsum(c, "complex64")
    net = conv.expectation_network(get_ham_gates([(q, "Z", 1.0) for q in range(30)]))
elif which == "c4":
    c = circuits.sycamore_rqc(int(sys.argv[3]) if len(sys.argv) > 3 else 14, seed=0)
    conv = QiboCircuitToEinsum(c, "complex64")
    net = conv.amplitude_network("0" * 53)
t = time.time()
info = find_path(net.modes, net.out, samples=samples, seed=0, target_log2_width=int(sys.argv[4]) if len(sys.argv) > 4 else 30)
print("plan %.1fs" % (time.time() - t), info.method, "log10flops %.2f W %d slices %d" % (math.log10(info.flops), info.log2_width, info.num_slices))
low = executor.lower(net.modes, net.out, info, "complex64")
BW, PK = 6.4e12, 60e12
groups = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for nd in low.nodes:
    p = nd.plan
    mult = info.num_slices if nd.depends else 1
    inten = p.flops / p.bytes
    cls = "tiny" if p.bytes < 1 << 20 else ("mem" if inten < PK / BW else "compute")
    key = (cls, p.kernel, (p.tmb, p.tnb, p.tkb), "slice" if nd.depends else "hoist")
    g = groups[key]
    g[0] += 1; g[1] += p.flops * mult; g[2] += p.bytes * mult
    g[3] += max(p.bytes / BW, p.flops / PK, 3e-6) * mult
tot = sum(g[3] for g in groups.values())
for key, g in sorted(groups.items(), key=lambda kv: -kv[1][3]):
    print(key, "n=%d flops=%.2e bytes=%.2e ideal_s=%.4f (%.0f%%)" % (g[0], g[1], g[2], g[3], 100 * g[3] / tot))
print("ideal total %.4f s; launches/slice %d hoisted %d; arena MB hoist %.1f slice %.1f" % (tot, low.launches_per_slice, low.launches_hoisted, low.hoist_bytes / 2**20, low.slice_bytes / 2**20))
