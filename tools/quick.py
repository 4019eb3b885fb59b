"""Scratch driver: bounded BFS of one configuration through the Python mirror; prints the level table."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.load()
R, V, L = map(int, sys.argv[1:4])
secs = float(sys.argv[4])
stop = int(sys.argv[5]) if len(sys.argv) > 5 else 1
table = int(sys.argv[6]) if len(sys.argv) > 6 else 0
frontier = int(sys.argv[7]) if len(sys.argv) > 7 else 0
import os
mc = pkg.ModelChecker.from_constants(R, V, L)
t0 = time.time()
res = mc.check(max_depth=int(os.environ.get('MAXDEPTH', '0')), max_seconds=secs, keep_trace=True, stop_on_violation=bool(stop), table_capacity=table, frontier_capacity=frontier)
print(json.dumps(dict(cfg=[R,V,L], rc=res.rc, distinct=res.distinct, generated=res.generated, depth=res.depth, complete=res.complete,
   wall=time.time()-t0, secs=res.seconds_total, kern=res.seconds_kernels, rate=res.distinct/res.seconds_total, krate=res.distinct/max(res.seconds_kernels,1e-9),
   g=res.generated/max(res.distinct,1), probes_per_gen=res.probe_total/max(res.generated,1), ties=res.h2_ties, coll=res.fp_collisions,
   viol_level=res.violation_level, table=res.table_capacity, frontier=res.frontier_capacity, trace=[a for a,_ in res.trace])))
n=len(res.level_sizes) if not os.environ.get('QUIET') else 0
for i in range(n):
    print(i+1, res.level_sizes[i], res.level_generated[i], round(res.level_ms[i],3), round(res.level_sizes[i]/max(res.level_ms[i],1e-6)/1e3,2), "M new/s")
if res.trace and len(sys.argv) > 8:
    open(sys.argv[8], "w").write(mc.dump_trace_tlc(res.trace))
