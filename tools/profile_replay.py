"""cProfile of a strategy-trace replay on the GPU path (where does a proposal's wall-clock go on the host side?).
usage: profile_replay.py trace_sop.npz [host|device] [max_calls]"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pysot_b200 as pb
from trace_replay import GpuBackend, replay
name = sys.argv[1] if len(sys.argv) > 1 else "trace_sop.npz"
mode = sys.argv[2] if len(sys.argv) > 2 else "device"
mx = int(sys.argv[3]) if len(sys.argv) > 3 else None
pb.set_candidate_generator(mode, seed=1)
replay(name, GpuBackend(), max_calls=30)
pr = cProfile.Profile()
pr.enable()
replay(name, GpuBackend(), max_calls=mx)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
