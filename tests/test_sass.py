"""Static checks of the SASS of the shipped library (tools/sass_check.py): they need cuobjdump and
the built liblapf.so, no GPU.  They guard the two places where the sampler leaves what ptxas
checks by itself: the split tcgen05.ld / tcgen05.wait::ld pair, and warp convergence before the
`.sync.aligned` TMEM loads."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def kernels():
    import sass_check
    from olpefit_b200 import _lib
    _lib.load()                                   # builds the library if the sources are newer
    fns = sass_check.functions(_lib.lib_path())
    batch = {k: v for k, v in fns.items() if "gibbs_batch_kernel" in k}
    assert len(batch) == 6, sorted(batch)
    return sass_check, fns, batch


def test_native_instructions_are_present(kernels):
    """The sampler kernels really use TMEM loads/stores, TMA bulk copies and packed FP32 -- and no
    tensor-core MMA: the path is not a contraction."""
    sass_check, fns, batch = kernels
    for name, ins in batch.items():
        ops = [x.op for x in ins]
        assert any(o.startswith("LDTM") for o in ops) and any(o.startswith("STTM") for o in ops), name
        assert sum(o == "UBLKCP.S.G" for o in ops) in (2, 4), name      # two planes per stamp slot (two slots up to 64 pixels)
        assert sum(o == "FFMA2" for o in ops) > 100 and "MUFU.EX2" in ops, name
        assert not any("MMA" in o for o in ops), name
    prep = [v for k, v in fns.items() if "frame_prep_tma_kernel" in k]
    assert prep and any(x.op.startswith("UTMALDG") for x in prep[0])     # the tensor-map cut-out


def test_every_tmem_load_is_waited_for_before_its_registers_are_touched(kernels):
    sass_check, fns, batch = kernels
    for name, ins in batch.items():
        bad, gaps = sass_check.check_ldtm_waits(ins)
        assert not bad, (name, bad)
        assert gaps and min(gaps) >= 1, name


def test_every_tmem_load_follows_an_unconditional_convergence_point(kernels):
    """WARPSYNC.ALL (not just the BRA.DIV run-time check) between the warp's last possible divergence
    and every `.sync.aligned` TMEM load: the one structural difference between the builds whose
    sampler mis-evaluated the update after a recorded update and every build that works."""
    sass_check, fns, batch = kernels
    for name, ins in batch.items():
        assert not sass_check.check_ldtm_convergence(ins), (name, sass_check.check_ldtm_convergence(ins))
