"""Host-side check of the tensor-core operand layouts of pm_tiles_imma: a NumPy emulation of the mma.sync.m16n8k32
fragment ownership (tests/imma_layout_emulation.py) must reproduce the direct correlation for the shipped ("perm") and
the candidate ("std", -DSID_IMMA_KSTD) K mapping, and reports the shared-memory wavefronts of the A loads."""
from tests import imma_layout_emulation as emu


def test_both_layouts_reproduce_direct_correlation():
    for layout in ("perm", "std"):
        assert emu.emulate(layout, s=8, R=12, seed=1)
        assert emu.emulate(layout, s=17, R=26, seed=2)       # two row blocks, two column groups, ragged edges


def test_bank_conflicts_of_a_loads():
    assert emu.a_load_wavefronts("perm", 24) == 2             # shipped: pitch 24 words on cfg2, always 2-way
    assert all(emu.a_load_wavefronts("perm", w) >= 2 for w in range(8, 72, 4))      # only even banks are touched
    assert all(emu.a_load_wavefronts("perm", w) == 2 for w in range(8, 72, 16))     # the shipped pitches, == 8 (mod 16)
    assert all(emu.a_load_wavefronts("std", w) == 1 for w in range(4, 72, 8))      # pitch == 4 (mod 8): conflict-free
    assert emu.a_load_wavefronts("std", 24) == 2
