"""Work decomposition of the pipelined kernels, checked on the CPU through the C ABI.

``b200_selftest_pipe_geometry`` runs the same inline functions the CUDA kernels use (copy shares,
chunk arrival counts, reduce work items, ring placement) on the host and verifies the invariants
the kernels rely on.  This covers message shapes the GPU tests cannot enumerate."""
import itertools

import pytest

from ray_b200 import _native as N

MiB = 1 << 20
TILE = 32 << 10


def _check(lib, nbytes, chunk, copy, world, red, ring=0):
    rc = lib.b200_selftest_pipe_geometry(nbytes, chunk, copy, world, red, ring)
    assert rc == 0, (nbytes, chunk, copy, world, red, ring, N.last_error())


def test_copy_shares_and_reduce_items_tile_every_message(native_lib):
    sizes = [16, 32, 16 * 1023, TILE - 16, TILE, TILE + 16, MiB - 16, MiB, MiB + 16, 3 * MiB + 16 * 77, 5 * MiB,
             17 * MiB + 48, 64 * MiB]
    for nbytes, (chunk, copy), world, red in itertools.product(
            sizes, [(MiB, 1), (MiB, 8), (MiB, 32), (2 * MiB, 16), (4 * MiB, 16), (8 * MiB, 16), (3 * MiB, 2)],
            [2, 3, 4, 5, 8], [1, 3, 32, 48, 64]):
        _check(native_lib, nbytes, chunk, copy, world, red)


def test_ring_placement_for_messages_larger_than_the_slot(native_lib):
    for nbytes in (41 * MiB + 16, 97 * MiB + 48, 256 * MiB, 1024 * MiB):
        for chunk, copy, ring in ((MiB, 16, 40), (4 * MiB, 16, 10), (8 * MiB, 16, 32), (8 * MiB, 8, 4)):
            for world in (3, 4, 8):
                _check(native_lib, nbytes, chunk, copy, world, 32, ring)


def test_property_random_shapes(native_lib):
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=300, deadline=None)
    @given(units=st.integers(1, (96 * MiB) // 16), chunk_mib=st.integers(1, 8), copy_log=st.integers(0, 5),
           world=st.integers(2, 8), red=st.integers(1, 96), ring=st.sampled_from([0, 4, 7, 32]))
    def prop(units, chunk_mib, copy_log, world, red, ring):
        _check(native_lib, units * 16, chunk_mib * MiB, 1 << copy_log, world, red, ring)

    prop()


def test_selftest_rejects_inconsistent_arguments(native_lib):
    assert native_lib.b200_selftest_pipe_geometry(15, MiB, 8, 4, 32, 0) == N.ERR_INVALID      # not 16-byte units
    assert native_lib.b200_selftest_pipe_geometry(MiB, MiB + TILE, 8, 4, 32, 0) == N.ERR_INVALID  # C % (G*tile)
    assert native_lib.b200_selftest_pipe_geometry(MiB, MiB, 8, 9, 32, 0) == N.ERR_INVALID      # world > 8
