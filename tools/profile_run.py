"""cProfile of gat_b200.run on one BASELINE configuration (host-side hot spots) -- profiling aid

    python tools/profile_run.py ns
"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import gat_b200  # noqa: E402
from gat_b200 import engine as Engine, synthetic  # noqa: E402
from tools.baseline_configs import CONFIGS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ns"
nseg, nanno, iso, counter, S = CONFIGS[name]
segments, annotations, workspaces, isochores = synthetic.make(nseg, nanno, 20000, isochores=iso)
workspace = synthetic.prepare(segments, annotations, workspaces, isochores)


def go():
    Engine.seed(1)
    res = gat_b200.run(segments, annotations, workspace, Engine.SamplerAnnotator(bucket_size=1, nbuckets=100000),
                       [Engine.COUNTER_CLASSES[counter]()], Engine.UnconditionalWorkspace(), num_samples=S)
    torch.cuda.synchronize()
    return res


go()                                    # warm-up: context, allocator pools
pr = cProfile.Profile()
pr.enable()
go()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
