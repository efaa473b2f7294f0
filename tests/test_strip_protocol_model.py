"""Model check of the row-strip exchange protocol (csrc/td_peer.cu td_push_regions + the in-kernel wait of
td_blend_multidiffusion_rows; parallel.StripExchange), no GPU: every rank is a little state machine whose micro-operations
(one CTA's data stores, one CTA's release add, one poll of a wait, one read of a halo) are interleaved at random.

Checked for many geometries / world sizes / schedules:
  * no deadlock: every rank finishes every step;
  * a blend reads, from every sender, exactly the halo of ITS step (never a stale one, never one overwritten by the
    sender's next step), and the next scatter reads the latent halo of its step;
  * the senders of a rank form a contiguous rank range (the wait polls slots [first, first + count)).
Per step and rank:  push_tiles (no counter of this rank touched)  ->  blend (wait: tile slots >= tile_expect)  ->
push_x (x_expect += P; push; wait: x slots >= x_expect; tile_expect += P).  tile_expect starts at P."""
import random

import pytest

from multidiffusion_upscaler_for_automatic1111_b200 import engine, parallel

P = 4          # CTAs per push in the model (TD_PUSH_CTAS = 32 on the device: only the count differs)


def _shards(W, H, tile, overlap, world):
    g = engine.make_grid(W, H, tile, tile, overlap, 4)
    return [parallel.StripShard(list(g.ys[:g.rows]), g.cols, g.tile_h, g.H, r, world) for r in range(world)]


def _rank_program(r, shards, state, steps):
    """Generator of micro-operations of rank r; yields after each one (True = progressed, False = polled and must retry)."""
    sh = shards[r]
    tile_targets = sorted({q for (_, q, _, _) in sh.halo_out()})
    tile_senders = sorted({p for (_, p, _, _) in sh.halo_in()})
    x_targets = sorted({p for (p, _, _) in sh.x_out()})
    x_senders = sorted({q for (q, _, _) in sh.x_in()})
    for group in (tile_senders, x_senders):
        assert group == list(range(group[0], group[-1] + 1)) if group else True, f"rank {r}: senders {group} are not a contiguous range"
    for s in range(1, steps + 1):
        # scatter: reads the latent halo rows published at the end of step s - 1
        for q in x_senders:
            assert state["x_data"][r][q] == [s - 1] * P, f"rank {r} step {s}: scatter saw latent halo {state['x_data'][r][q]} from {q}"
            yield True
        # push_tiles: every CTA stores its share, then adds 1 (release) to each target's slot
        for cta in range(P):
            for q in tile_targets:
                state["tile_data"][q][r][cta] = s
                yield True
            for q in tile_targets:
                state["tile_slot"][q][r] += 1
                yield True
        # blend: the halo CTAs poll until every sender's slot has reached this rank's expect counter, then read
        want = state["tile_expect"][r]
        while any(state["tile_slot"][r][p] < want for p in tile_senders):
            yield False
        for p in tile_senders:
            yield True          # (other ranks may run between the wait and the read)
            assert state["tile_data"][r][p] == [s] * P, f"rank {r} step {s}: blend read halo {state['tile_data'][r][p]} of rank {p}"
        # push_x: bump own expect, push + signal, wait for own senders, then advance the tile set's expect counter
        state["x_expect"][r] += P
        yield True
        for cta in range(P):
            for q in x_targets:
                state["x_data"][q][r][cta] = s
                yield True
            for q in x_targets:
                state["x_slot"][q][r] += 1
                yield True
        while any(state["x_slot"][r][q] < state["x_expect"][r] for q in x_senders):
            yield False
        state["tile_expect"][r] += P
        yield True
    state["done"][r] = True


def _run(shards, steps, rng):
    world = len(shards)
    zeros = lambda: [[0] * world for _ in range(world)]
    state = {"tile_slot": zeros(), "x_slot": zeros(), "tile_expect": [P] * world, "x_expect": [0] * world,
             "tile_data": [[[0] * P for _ in range(world)] for _ in range(world)],
             "x_data": [[[0] * P for _ in range(world)] for _ in range(world)], "done": [False] * world}
    progs = [_rank_program(r, shards, state, steps) for r in range(world)]
    live = list(range(world))
    stalled = 0
    while live:
        r = rng.choice(live)
        burst = rng.choice((1, 1, 2, 5, 40))          # short and long time slices
        for _ in range(burst):
            try:
                progressed = next(progs[r])
            except StopIteration:
                live.remove(r)
                break
            stalled = 0 if progressed else stalled + 1
            if not progressed:
                break
        assert stalled < 20000 * world, f"deadlock: ranks {live} all polling, state {state['tile_slot']} / {state['tile_expect']}"
    assert all(state["done"])


@pytest.mark.parametrize("geom", [(512, 512, 96, 48), (512, 1024, 96, 48), (256, 640, 64, 16), (128, 320, 96, 8)])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_strip_protocol_under_random_schedules(geom, world):
    W, H, tile, overlap = geom
    shards = _shards(W, H, tile, overlap, world)
    for seed in range(12):
        _run(shards, steps=4, rng=random.Random(1000 * world + seed))


def test_a_stale_expect_counter_would_be_caught():
    """The model is sharp: if the tile expect counter were advanced by the tile push itself (i.e. read by a concurrent blend
    before the bump), a blend could run on the previous step's halo -- emulate by starting tile_expect one push behind."""
    shards = _shards(512, 512, 96, 48, 2)
    global P
    with pytest.raises(AssertionError):
        for seed in range(200):
            rng = random.Random(seed)
            world = 2
            zeros = lambda: [[0] * world for _ in range(world)]
            state = {"tile_slot": zeros(), "x_slot": zeros(), "tile_expect": [0] * world, "x_expect": [0] * world,
                     "tile_data": [[[0] * P for _ in range(world)] for _ in range(world)],
                     "x_data": [[[0] * P for _ in range(world)] for _ in range(world)], "done": [False] * world}
            progs = [_rank_program(r, shards, state, 3) for r in range(world)]
            live = [0, 1]
            while live:
                r = rng.choice(live)
                try:
                    next(progs[r])
                except StopIteration:
                    live.remove(r)
