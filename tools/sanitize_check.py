"""Small workload for compute-sanitizer (memcheck / racecheck): all three decoder kinds on tiny batches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, load_golden_bp4
from slidingwindowdecoder_b200 import bpgdg_decoder, bpgd_decoder, osd_window, bp4_osd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for name, cls in (("c2_w1_gdg_mt1", bpgdg_decoder), ("c2_w1_gdg_mt0", bpgdg_decoder), ("c1_bpgd", bpgd_decoder),
                  ("c2_w1_osdw_cs10", osd_window), ("c5_w0_osdw_cs10", osd_window)):
    g = load_golden(name)
    d = cls(g["mat"], channel_probs=g["priors"], **g["kwargs"])
    corr, conv = d.decode_batch(g["synd"][:n])
    print(name, "ok", bool(np.array_equal(corr, g["dec"][:n]) or name.endswith("mt1")), int(conv.sum()))

# a [[144,12,12]] window (n = 1728): the radix-select / partial-sort path of sort_reset_kernel
g = load_golden("c3_w5_gdg_mt1")
d = bpgdg_decoder(g["mat"], channel_probs=g["priors"], **g["kwargs"])
k = max(4, n // 6)
corr, conv = d.decode_batch(g["synd"][:k])
print("c3_w5_gdg_mt1", "ok", int((corr == g["dec"][:k]).all(axis=1).sum()), "of", k, int(conv.sum()))
# bp4_osd decode + camel_decode on the pair with an all-ones last column
g = load_golden_bp4("c1_bp4_camel_tied")
d = bp4_osd(g["hx"], g["hz"], channel_probs_x=g["px"], channel_probs_y=g["py"], channel_probs_z=g["pz"], **g["kwargs"])
out = d.camel_decode_batch(g["synd_x"][:n], g["synd_z"][:n])
print("c1_bp4_camel_tied", "ok", bool(np.array_equal(out["converge"], g["conv"][:n])), int(out["converge"].sum()))
out = d.decode_batch(g["synd_x"][:n], g["synd_z"][:n])
print("c1_bp4 decode (tied)", "ok", int(out["converge"].sum()))

# round 2: the large-graph paths forced on small windows (HBM-streamed BP with its cp.async rings, big radix select, large-T OSD
# layout), the bit-packed entry point, and a few shots of the un-windowed [[144,12,12]] DEM itself
from slidingwindowdecoder_b200.decoders import pack_bits, unpack_bits
for env in ("SWD_FORCE_STREAM", "SWD_FORCE_BIG_SORT", "SWD_FORCE_BIG_OSD"):
    os.environ[env] = "1"
g = load_golden("c2_w1_osdw_cs10")
d = osd_window(g["mat"], channel_probs=g["priors"], **g["kwargs"])
pc, conv = d.decode_batch_packed(pack_bits(g["synd"][:n]))
print("c2_w1_osdw_cs10 forced stream / big sort / big osd, packed", "ok", bool(np.array_equal(unpack_bits(pc, d.n), g["dec"][:n])), int(conv.sum()))
g = load_golden("c2_w1_gdg_mt1")
d = bpgdg_decoder(g["mat"], channel_probs=g["priors"], **g["kwargs"])
corr, conv = d.decode_batch(g["synd"][:n])
print("c2_w1_gdg_mt1 forced stream / big sort", "ok", int(conv.sum()))
for env in ("SWD_FORCE_STREAM", "SWD_FORCE_BIG_SORT", "SWD_FORCE_BIG_OSD"):
    del os.environ[env]
if os.environ.get("SANITIZE_G144", "1") == "1":
    g = load_golden("g144_osdw_cs10")
    d = osd_window(g["mat"], channel_probs=g["priors"], **g["kwargs"])
    corr, conv = d.decode_batch(g["synd"][:3])
    print("g144_osdw_cs10", "ok", bool(np.array_equal(corr, g["dec"][:3])), int(conv.sum()))
