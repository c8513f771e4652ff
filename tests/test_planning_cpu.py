"""Planning helpers of the host side — pure integer logic that decides how the kernels are launched: the line
convolution's (channels per pass, rows per work item) against the 512 TMEM columns, the frame ranges of the
temporally sharded VAE (reference chunking: wan_vae.py:520-575; the sharding itself is new), the head groups of the
sequence-parallel head exchange.  Properties, checked over every width / clip length / rank count that can occur."""
import pytest

from videocof_b200 import vae


def _nacc(n_tile):
    return min(5, 512 // ((n_tile + 31) // 32 * 32))


@pytest.mark.parametrize("n_total", [16, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 80, 144])
@pytest.mark.parametrize("fused", [False, True])
def test_lines_plan_fits_tensor_memory(n_total, fused, monkeypatch):
    monkeypatch.delenv("VCOF_CONV_SPARE", raising=False)
    n_tile, rows = vae._lines_plan(n_total, fused)
    assert n_tile == n_total if n_total <= 128 else n_total % n_tile == 0      # whole channel passes
    assert n_tile <= 128 and n_tile % 16 == 0
    nacc = _nacc(n_tile)
    assert 1 <= rows <= min(4, nacc)
    assert nacc * ((n_tile + 31) // 32 * 32) <= 512                            # the accumulator ring fits TMEM
    if fused and nacc >= 3:
        assert nacc - rows >= 2      # two spare accumulators absorb the fused epilogue's burst
    if not fused:
        assert rows == min(4, nacc)  # plain layers: as many rows as fit, the weights of a phase serve more rows


def test_lines_plan_vae_layers(monkeypatch):
    """The widths of the Wan VAE (96 / 192 / 384 and the 16-padded heads) and what they resolve to."""
    monkeypatch.delenv("VCOF_CONV_SPARE", raising=False)
    assert vae._lines_plan(96, True) == (96, 3) and vae._lines_plan(96, False) == (96, 4)
    assert vae._lines_plan(192) == (96, 4)
    assert vae._lines_plan(384) == (128, 4)
    assert vae._lines_plan(16) == (16, 4)
    monkeypatch.setenv("VCOF_CONV_SPARE", "1")
    assert vae._lines_plan(96, True) == (96, 4) and vae._lines_plan(384) == (128, 3)


class _Shard(vae.TimeShard):
    def __init__(self, world, rank=0):          # no process group: plan() is pure arithmetic on world / rank
        self.world, self.rank, self.active = world, rank, world


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_time_shard_plan_properties(world):
    for f in range(1, 90):
        ts = _Shard(world)
        counts = ts.plan(f)
        if min(world, f // 2) < 2:
            assert counts is None                # one active rank: run un-sharded
            continue
        assert len(counts) == world and sum(counts) == f
        active = [c for c in counts if c]
        assert ts.active == len(active) == min(world, f // 2)
        assert counts[:ts.active] == active and all(c == 0 for c in counts[ts.active:])   # idle ranks trail
        assert min(active) >= 2                  # every active rank can fill its right neighbour's 2-frame halo
        assert max(active) - min(active) <= 1 and active == sorted(active, reverse=True)   # balanced, extras first


def test_time_shard_plan_c2_and_c5_clips():
    """The clips of BASELINE configs on 8 GPUs: C2 encodes 10 latent frames (5 active ranks), decodes 10 (edit) and 1
    (ground: un-sharded); C5 works on 40 / 81-frame clips (all 8 ranks)."""
    assert _Shard(8).plan(10) == [2, 2, 2, 2, 2, 0, 0, 0]
    assert _Shard(8).plan(1) is None
    assert _Shard(8).plan(21) == [3, 3, 3, 3, 3, 2, 2, 2]
    assert _Shard(8).plan(40) == [5] * 8
    assert _Shard(4).plan(21) == [6, 5, 5, 5]


@pytest.mark.parametrize("local_heads", range(1, 41))
def test_head_groups_cover_the_local_heads(local_heads, monkeypatch):
    from videocof_b200.dist import SequenceParallel
    sp = SequenceParallel.__new__(SequenceParallel)
    monkeypatch.delenv("VCOF_SP_SPLIT", raising=False)
    g = sp._head_groups(local_heads)
    assert sum(g) == local_heads and all(x > 0 for x in g)
    assert len(g) == (1 if local_heads < 2 else 2) and g == sorted(g, reverse=True)   # the larger group goes first
    monkeypatch.setenv("VCOF_SP_SPLIT", "0")
    assert sp._head_groups(local_heads) == [local_heads]
