import sys, json, torch
sys.path.insert(0, '.')
import bench
from pairnet_b200 import _native as nat
lib = nat.load()
for opt in (1, 2, 1, 2):
    lib.pn_set_option(nat.PN_OPT_PPN_EPI2, opt)
    r = bench.ppn_microbench("cuda", bench.peaks())
    print(opt, [(x["N"], round(x["frac_of_hbm_peak"], 3), round(x["bf16"]["frac_of_hbm_peak"], 3)) for x in r], flush=True)
lib.pn_set_option(nat.PN_OPT_PPN_EPI2, 1)
