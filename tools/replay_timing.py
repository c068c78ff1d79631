"""Wall-clock of BASELINE's strategy-level configs (#1, #2, #5) on the GPU path, beside the reference's own time for
the same calls (recorded by tools/make_strategy_traces.py on the build container's CPU; it is a different host than
the GPU box, so the ratio is indicative only).  One "call" = what a strategy does per proposal: lazy refit of the
surrogate + candidate generation + weighted_distance_merit."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pysot_b200 as pb
from trace_replay import GpuBackend, OracleBackend, replay
from conftest import GOLDEN

out = []
for name, label in (("trace_dycors.npz", "#1 DYCORS Ackley-10 serial, 500 evals"),
                    ("trace_srbf.npz", "#2 SRBF Levy-30 async x8, 2000 evals"),
                    ("trace_sop.npz", "#5 SOP Rastrigin-20 8 centers, 3000 evals")):
    g = np.load(os.path.join(GOLDEN, name))
    ref = float(g["ref_call_seconds"].sum()) if "ref_call_seconds" in g.files else float("nan")
    row = {"config": label, "calls": int(len(g["n"])), "reference_cpu_seconds": ref}
    for mode in ("host", "device"):
        pb.set_candidate_generator(mode, seed=1)
        replay(name, GpuBackend(), max_calls=20)  # warm-up (library load, buffers)
        tm = {}
        ncalls, mism, ties, err = replay(name, GpuBackend(), timer=tm)
        row["gpu_seconds_%s_candidates" % mode] = tm["seconds"]
        if mode == "host":
            row["mismatches"] = len(mism)
    pb.set_candidate_generator("host")
    if "--no-oracle" not in sys.argv:  # the oracle (bit-identical to the reference, incremental LU included) on THIS host
        tm = {}
        replay(name, OracleBackend(), timer=tm)
        row["oracle_cpu_seconds_this_host"] = tm["seconds"]
    out.append(row)
    print(json.dumps(row), flush=True)
