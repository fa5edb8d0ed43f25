"""Index logic of the register bitonic network of the symbolic phase (csrc/csr.cu: warp_bitonic_regs), restated for the CPU.

The device code cannot run without a GPU; what CAN be checked here is the part that is easy to get wrong -- which compare-exchange
steps stay inside a lane, which go through a shuffle with lane ``l ^ (stride / M)``, and the direction of every step -- by running
the same index arithmetic over a simulated warp.  The GPU tests (CSR pattern bit-identical to scipy on every fixture) check the
kernel itself."""
import random

import pytest


def warp_bitonic_regs(keys, M):
    """keys: 32 * M values, lane l holds keys[l*M : l*M+M]; returns the network's output in the same layout."""
    k = [[keys[l * M + r] for r in range(M)] for l in range(32)]
    size = 2
    while size <= 32 * M:
        stride = size >> 1
        while stride > 0:
            if stride >= M:                                    # partner in lane l ^ (stride / M), same slot
                new = [row[:] for row in k]
                for lane in range(32):
                    i0 = lane * M
                    take_min = ((i0 & size) == 0) == ((i0 & stride) == 0)
                    for r in range(M):
                        o = k[lane ^ (stride // M)][r]
                        new[lane][r] = k[lane][r] if ((k[lane][r] < o) == take_min) else o
                k = new
            else:                                              # slots r and r | stride of the same lane
                for lane in range(32):
                    i0 = lane * M
                    for r in range(M):
                        if (r & stride) == 0:
                            r2 = r | stride
                            up = ((i0 & size) == 0) if size >= M else ((r & size) == 0)
                            a, c = k[lane][r], k[lane][r2]
                            if (a > c) == up:
                                k[lane][r], k[lane][r2] = c, a
            stride >>= 1
        size <<= 1
    return [k[l][r] for l in range(32) for r in range(M)]


@pytest.mark.parametrize("M", [2, 4, 8, 16])
def test_network_sorts_padded_rows(M):
    rng = random.Random(M)
    pad = 2 ** 64 - 1
    for _ in range(60):
        L = rng.randint(1, 32 * M)
        # keys as the kernel builds them: (column << 32) | source, sources unique
        keys = [(rng.randrange(1 << 20) << 32) | s for s in rng.sample(range(1 << 31), L)] + [pad] * (32 * M - L)
        assert warp_bitonic_regs(keys, M) == sorted(keys)
