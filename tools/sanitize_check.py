"""Small workload for compute-sanitizer (memcheck / racecheck): all three decoder kinds on tiny batches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from slidingwindowdecoder_b200 import bpgdg_decoder, bpgd_decoder, osd_window

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for name, cls in (("c2_w1_gdg_mt1", bpgdg_decoder), ("c2_w1_gdg_mt0", bpgdg_decoder), ("c1_bpgd", bpgd_decoder),
                  ("c2_w1_osdw_cs10", osd_window), ("c5_w0_osdw_cs10", osd_window)):
    g = load_golden(name)
    d = cls(g["mat"], channel_probs=g["priors"], **g["kwargs"])
    corr, conv = d.decode_batch(g["synd"][:n])
    print(name, "ok", bool(np.array_equal(corr, g["dec"][:n]) or name.endswith("mt1")), int(conv.sum()))
