"""One device-resident step of the bench workload (32768 n=12 / 4-fold partitions) for ncu captures:
    AMBI_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 26 -c 2 -o gpurun_out/prof python tools/prof_target.py 32768
(-s 26: the first 12 k_expand launches are the greedy dive, the next 12 the beam; 26 and 27 are levels 2 and 3 of the exact pass)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
import bench
from ambigram_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
problems = bench.make_problems(n, 0)
ctx = capi.Context(device=0)
b = ctx.upload(problems)
b.run(); ctx.sync()
sols = b.fetch()
print("ok", sum(s.candidates_scored for s in sols))
b.free(); ctx.close()
