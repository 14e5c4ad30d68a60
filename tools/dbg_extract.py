import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np
from pf_loader import pfb
import oracle as O
capi = pfb.capi
p = pfb.synth.config("cfg2")
ex = capi.Extractor(num_lines=64, max_points=131072)
for f in (3, 7, 33):
    s = pfb.synth.scan(p, f)
    for rep in range(2):
        edge, surf, label = ex.run(s)
        ref = O.extract(s, order=1)
        bad = np.nonzero(label != ref["label"])[0]
        print("frame", f, "rep", rep, "label mismatches", len(bad), "edges", len(edge), len(ref["edge_idx"]), "surf", len(surf), len(ref["surf_idx"]))
        if len(bad):
            ring = ref["ring"]
            print("  idx", bad[:20], "rings", ring[bad[:20]], "got", label[bad[:20]], "want", ref["label"][bad[:20]])
            # position in ring
            for b in bad[:6]:
                r = ring[b]; pos = np.count_nonzero(ring[:b] == r); nr = np.count_nonzero(ring == r)
                L = (nr - 10) // 6
                print("   ring", r, "pos", pos, "nr", nr, "L", L, "sector", (pos - 5) // L if L else -1, "rel", (pos - 5) - L * ((pos - 5) // L))
