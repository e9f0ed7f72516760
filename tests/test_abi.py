"""CPU-only: the C-ABI library loads and exports every symbol include/brutus_b200.h declares; the
product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from brutus_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_match_header(lib):
    from brutus_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "brutus_b200.h")).read()
    declared = set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None


def test_struct_layout_matches_header(lib):
    from brutus_b200 import _lib
    o = _lib.Options()
    lib.bf_default_options(C.byref(o))
    assert tuple(o.avlim) == (0., 20.) and tuple(o.av_gauss) == (0., 1e6)
    assert tuple(o.rvlim) == (1., 8.) and tuple(o.rv_gauss) == (3.32, 0.18)
    assert (o.ltol, o.ltol_subthresh, o.init_thresh, o.wt_thresh, o.select_slack) == (3e-2, 1e-2, 5e-3, 1e-3, 0.5)
    assert (o.dim_prior, o.max_iter, o.apply_parallax_clip) == (1, 0, 1)
    assert C.sizeof(_lib.Options) == 8 * 8 + 5 * 8 + 4 * 4
    assert C.sizeof(_lib.Stats) == 4 * 8 + 10 * 8 + 3 * 8 + 5 * 8
    assert C.sizeof(_lib.Records) == 8 + 8 + 4 + 4 + 8 + 8
    po = _lib.PostOptions()
    lib.bf_default_post_options(C.byref(po))
    assert (po.nmc_prior, po.ndraws, po.seed, po.use_gal_prior, po.star_base) == (50, 250, 0, 1, 0)
    # defaults of gal_lnprior (brutus/pdf.py:476-486) survive the struct round trip, last field included
    assert (po.gal.R_solar, po.gal.Z_thick, po.gal.f_halo, po.gal.min_sigma) == (8.2, 0.9, 0.005, 1.0)
    assert (po.gal.galcen_distance, po.gal.z_sun) == (8.122, 0.0208)
    assert not po.z_override and not po.u_override
    assert C.sizeof(_lib.PostOptions) == 4 + 4 + 8 + 4 + 4 + 8 + 8 + 30 * 8 + 2 * 8
    assert po.nsel_max == 0
    assert C.sizeof(_lib.Draws) == 10 * 8


def test_no_cpu_fallback(lib):
    from brutus_b200 import _lib
    if lib.bf_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.BrutusCudaError):
        _lib.Handle(0)
    import numpy as np
    from brutus_b200 import fitting, mock
    grid, _ = mock.make_grid(100, 5, seed=1)
    st = mock.make_stars(grid, 1, seed=2)
    with pytest.raises(_lib.BrutusCudaError):
        fitting.loglike(st["flux"][0], st["err"][0], st["mask"][0].copy(), grid)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "brutus_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), (f, "product code must not touch oracle/")


def test_ctypes_structs_match_the_header_as_compiled(tmp_path):
    """sizeof / offsetof of every struct of include/brutus_b200.h as a C compiler sees them == the ctypes mirror."""
    import subprocess
    from brutus_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "brutus_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu\n", sizeof(bf_options), sizeof(bf_stats), sizeof(bf_records), sizeof(bf_gal_params),
           sizeof(bf_post_options), sizeof(bf_draws));
    printf("%zu %zu %zu %zu %zu\n", offsetof(bf_options, dim_prior), offsetof(bf_stats, ms_post),
           offsetof(bf_post_options, star_base), offsetof(bf_post_options, gal), offsetof(bf_post_options, z_override));
    return 0;
}''')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [C.sizeof(t) for t in (_lib.Options, _lib.Stats, _lib.Records, _lib.GalParams, _lib.PostOptions, _lib.Draws)]
    offs = [_lib.Options.dim_prior.offset, _lib.Stats.ms_post.offset, _lib.PostOptions.star_base.offset,
            _lib.PostOptions.gal.offset, _lib.PostOptions.z_override.offset]
    assert [int(x) for x in out[:6]] == sizes
    assert [int(x) for x in out[6:]] == offs
