import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
for env in ('0', '1'):
    os.environ['T2O_NM_RESIDENT'] = env
    t0 = time.time(); r = bench.planner_gier_run(torch.device("cuda:0"), 5010); print("resident", env, r, "pairs/s %.1f" % (bench.GIER_M / r[0]), time.time() - t0)
