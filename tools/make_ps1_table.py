"""Build brutus_b200/data/ps1_mr_lf.npz from the reference's PanSTARRS r-band luminosity-function table
(brutus/PSMrLF_lnprior.dat: two columns, Mr and ln prior; data, not code -- `ps1_MrLF_lnprior`,
brutus/pdf.py:111-141, linearly interpolates / extrapolates it).  Run in the build container:

    python tools/make_ps1_table.py [/root/reference]
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
mr, lnp = np.loadtxt(os.path.join(ref, "brutus", "PSMrLF_lnprior.dat")).T
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brutus_b200", "data", "ps1_mr_lf.npz")
np.savez_compressed(out, Mr=mr, lnprior=lnp)
print("wrote", out, len(mr), "rows, Mr", mr.min(), "..", mr.max())
