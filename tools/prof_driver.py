"""Where does the host spend the step?  cProfile of one resident pass of driver.genotype on bench.py's configs[1] batch.
python tools/prof_driver.py [scale]"""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
import bench
from bayestyper_b200 import capi, driver

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lib = capi.load()
capi.check(lib.btg_init(0), lib)
dev = torch.device("cuda", 0)
opt = driver.Options(random_seed=20190401)
inp = bench.build_batch(lib, 0, scale, dev)
inp.make_resident(lib, opt)
for _ in range(2):
    driver.genotype(inp, opt, resident=True)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
driver.genotype(inp, opt, resident=True)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
