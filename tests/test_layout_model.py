"""CPU models of two layout / launch-shape claims the CUDA code relies on (no GPU needed).

1. csrc/kron_fused.cuh keeps ONE copy of the 64 x 64 factor matrix in shared memory with a row stride of 68 doubles and reads DMMA B-fragments
   from it both straight (row = c0 + lane / 4, col = k + lane % 4) and transposed (row = k + lane % 4, col = c0 + lane / 4).  An 8-byte
   shared-memory access is served per half-warp; it is conflict-free when the 16 lanes of a half-warp touch 16 distinct 8-byte banks
   (address in doubles mod 16).  The claim holds for any stride == 4 (mod 16) -- and fails for the unpadded stride 64.
2. csrc/ozaki.cuh launch_kgrad_planes_t deals the rows of a chunk to ny row groups so that the grid is one wave of resident CTAs; every row
   must be covered exactly once, a CTA's row count must not fall below the staging tile, and the row-group count must not exceed the slots."""
import math

import pytest


def banks(stride, transposed, c0=0, k=0):
    out = []
    for half in (0, 1):
        b = []
        for lane in range(16 * half, 16 * half + 16):
            fr, fk = lane >> 2, lane & 3
            addr = (k + fk) * stride + c0 + fr if transposed else (c0 + fr) * stride + k + fk
            b.append(addr % 16)
        out.append(b)
    return out


@pytest.mark.parametrize("transposed", [False, True])
def test_fragment_reads_are_conflict_free_at_stride_68(transposed):
    for c0 in range(0, 64, 8):
        for k in range(0, 64, 4):
            for half in banks(68, transposed, c0, k):
                assert len(set(half)) == 16


def test_unpadded_stride_would_conflict():
    assert any(len(set(h)) < 16 for h in banks(64, False)) and any(len(set(h)) < 16 for h in banks(64, True))


def balanced_shape(rows, Mp, resident, max_slots, tile=128, small_limit=16384, small_rows=16):
    """restatement of launch_kgrad_planes_t: returns (rows per block, row groups)"""
    gx = (Mp + 255) // 256
    rpb = small_rows if rows <= small_limit else tile
    if rows > small_limit and max_slots > 0:
        ny = max(1, min(resident // gx, max_slots))
        r = (rows + ny - 1) // ny
        if r >= tile:
            rpb = r
    return rpb, (rows + rpb - 1) // rpb


@pytest.mark.parametrize("rows", [512, 16384, 16896, 23552, 65536, 147456, 2097152])
@pytest.mark.parametrize("Mp", [64, 192, 512, 1024, 4096])
def test_balanced_launch_covers_every_row_once_within_the_slots(rows, Mp):
    resident = 148 * 6
    slots = max(math.ceil(rows / 128), math.ceil(min(rows, 16384) / 16))           # kg_slot_count
    rpb, ny = balanced_shape(rows, Mp, resident, slots)
    assert ny <= slots
    covered = sum(max(0, min(rpb, rows - y * rpb)) for y in range(ny))
    assert covered == rows and (ny - 1) * rpb < rows
    if rows > 16384:
        assert rpb >= 128
        gx = (Mp + 255) // 256
        if rows // 128 * gx > resident:                                             # more 128-row blocks than one wave: the grid becomes one wave
            assert ny * gx <= resident
