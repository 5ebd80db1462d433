"""A schedule-exploring model of conv_chain_kernel's synchronisation protocol
(tecogan-pytorch_b200/csrc/tg_chain_tcgen05.cu) -- CPU only, no GPU, no library call.

The kernel chains 64->64 convolutions inside one persistent launch; its correctness rests on a
protocol between five roles per CTA (TMA producer, MMA issuer, weight streamer, dependency checker,
two epilogue groups) and on per-tile progress flags between CTAs.  Every role below is a transcription
of the corresponding loop of the kernel as a Python generator; asynchronous hardware (TMA
completions, the in-order tensor pipe with its tcgen05.commit arrivals) is modelled by event queues
that a seeded random scheduler drains in arbitrary interleavings.  The model checks, for many
shapes (tiles per CTA from 1 up, 1..6 layers, in-place residual buffers like SRNet's) and schedules:

  * no deadlock: some role or event can always make progress until every tile of every layer is done;
  * mbarrier parity waits are sound: a waiter is never more than one phase behind its barrier
    (the kernel waits on parities, so a second completion would make the wait hang);
  * RAW on activations: when a tile's TMA box lands, every tile under its halo holds the previous
    layer's output -- not older (not yet produced) and not newer (overwritten in place too early);
  * the weight ring: when an MMA executes, the slot it reads holds this layer's tap.

It reproduces the two bugs that bring-up hit on the GPU (`lookahead_blocks=True`: the MMA issuer
blocking on the next tile's data before committing the current one; `gate_streamer=False`: the
streamer completing a weight barrier twice before the issuer's first wait), and it found a third one
that the GPU tests never triggered (`gate_early=False`: with one tile per CTA a delayed streamer
could find a hand-over barrier two phases ahead) -- see test_model_catches_the_protocol_bugs.
"""
import random

import pytest

STAGES, BUFS, SLOTS, WINDOW = 4, 8, 14, 9


class ProtocolError(AssertionError):
    pass


class MBar:
    """mbarrier with an arrival count; waits name the completion they are for (0-based) so that the
    model can flag a waiter that fell two phases behind (the kernel only knows the parity)."""

    def __init__(self, count=1):
        self.count, self.arrived, self.phase = count, 0, 0

    def arrive(self):
        self.arrived += 1
        if self.arrived == self.count:
            self.arrived = 0
            self.phase += 1

    def done(self, k, who=''):
        if self.phase >= k + 2:
            raise ProtocolError(f'{who}: barrier is {self.phase - k} phases past the awaited completion {k}')
        return self.phase == k + 1


class Cta:
    def __init__(self, sim, b):
        self.sim, self.b = sim, b
        s = sim
        self.n_my = (s.num_tiles - 1 - b) // s.G + 1
        self.total = s.L * self.n_my
        self.full = [MBar() for _ in range(STAGES)]
        self.empty = [MBar() for _ in range(STAGES)]
        self.tfull = [MBar() for _ in range(BUFS)]
        self.tempty = [MBar(1) for _ in range(BUFS)]      # one arrival per epilogue group (4 warps in the kernel)
        self.wfull = [MBar() for _ in range(9)]
        self.wfree = [MBar() for _ in range(9)]
        self.wstart = MBar()
        self.wearly = MBar()
        self.deps_ok = self.n_my                         # layer 0 has no dependencies
        self.stage_tile = [None] * STAGES                # what the A stage holds (seq) once landed
        self.slot = [None] * SLOTS                       # (layer, tap) each weight slot holds
        self.pipe = []                                   # in-order tensor pipe: ('mma', seq, tap) | ('commit', MBar)
        self.loads = []                                  # in-flight TMA / bulk loads: callables
        self.done_tiles = 0

    def tile(self, seq):
        l, k = divmod(seq, self.n_my)
        return l, k, self.b + k * self.sim.G

    # ------------------------------------------------------------------ roles (generators: yield = blocked)
    def producer(self):
        s = self.sim
        for seq in range(self.total):
            l, k, t = self.tile(seq)
            if l > 0:
                while self.deps_ok <= seq:
                    yield
            st, use = seq % STAGES, seq // STAGES
            if use > 0:
                while not self.empty[st].done(use - 1, 'producer/empty'):
                    yield
            s.check_halo_ready(l, t, 'TMA issue')         # what the checker promised
            self.loads.append(lambda st=st, seq=seq, l=l, t=t: self._land(st, seq, l, t))
            yield

    def _land(self, st, seq, l, t):
        self.sim.check_halo_ready(l, t, 'TMA landing')    # nobody overwrote the halo in the meantime
        self.stage_tile[st] = seq
        self.full[st].arrive()

    def mma(self):
        s = self.sim
        while not (self.full[0].done(0, 'mma/full')):
            yield
        for seq in range(self.total):
            l, k, t = self.tile(seq)
            first_k, last_k = k == 0, k == self.n_my - 1
            hand_over = last_k and l + 1 < s.L
            st, buf = seq % STAGES, seq % BUFS
            nst, nbuf = (seq + 1) % STAGES, (seq + 1) % BUFS
            if first_k:
                for tap in range(9):
                    while not self.wfull[tap].done(l, 'mma/wfull'):
                        yield
                self.wstart.arrive()
            if hand_over and s.gate_early:
                # the streamer must have consumed the PREVIOUS completion of wfree[4..8] (its waits for
                # the next layer's taps 0-4) before this layer's hand-over completes them again
                while not self.wearly.done(l, 'mma/wearly'):
                    yield
            next_ready = False
            for part in range(2):
                for tap in (range(0, 4) if part == 0 else range(4, 9)):
                    self.pipe.append(('mma', seq, tap))
                if hand_over:
                    for tap in (range(0, 4) if part == 0 else range(4, 9)):
                        self.pipe.append(('commit', self.wfree[tap]))
                if part == 1:
                    self.pipe.append(('commit', self.empty[st]))
                    self.pipe.append(('commit', self.tfull[buf]))
                if part == 0 and seq + 1 < self.total:
                    def ready():
                        a = (seq + 1) < BUFS or self.tempty[nbuf].done((seq + 1) // BUFS - 1, 'mma/tempty')
                        return a and self.full[nst].done((seq + 1) // STAGES, 'mma/full')
                    if s.lookahead_blocks:                # the bring-up bug: a blocking wait here
                        while not ready():
                            yield
                    next_ready = ready()
                yield
            if seq + 1 < self.total and not next_ready:
                while not ((seq + 1) < BUFS or self.tempty[nbuf].done((seq + 1) // BUFS - 1, 'mma/tempty')):
                    yield
                while not self.full[nst].done((seq + 1) // STAGES, 'mma/full'):
                    yield

    def streamer(self):
        s = self.sim
        for tap in range(9):                              # layer 0, before anything else
            self.loads.append(lambda tap=tap: self._wland(0, tap))
        for l in range(1, s.L):
            if s.gate_streamer:
                while not self.wstart.done(l - 1, 'streamer/wstart'):
                    yield
            for tap in range(9):
                gi = 9 * l + tap
                if gi >= SLOTS:
                    pl, pt = divmod(gi - SLOTS, 9)
                    while not self.wfree[pt].done(pl, 'streamer/wfree'):
                        yield
                self.loads.append(lambda l=l, tap=tap: self._wland(l, tap))
                if tap == SLOTS - 9 - 1:
                    self.wearly.arrive()                  # the waits that refer to layer l-2 are behind us
                yield

    def _wland(self, l, tap):
        self.slot[(9 * l + tap) % SLOTS] = (l, tap)
        self.wfull[tap].arrive()

    def checker(self):
        s = self.sim
        head = self.n_my
        while head < self.total:
            p = 0
            while p < WINDOW and head + p < self.total:
                l, k, t = self.tile(head + p)
                if not all(s.flag[u] >= l for u in s.halo(t)):
                    break
                p += 1
            if p:
                head += p
                self.deps_ok = head
            yield

    def epilogue(self, grp):
        s = self.sim
        pending = None
        for seq in range(grp, self.total, 2):
            l, k, t = self.tile(seq)
            buf = seq % BUFS
            if not self.tfull[buf].done(seq // BUFS, 'epilogue/tfull'):
                if pending is not None:                   # next accumulator not ready: publish at once
                    s.publish(*pending)
                    pending = None
                while not self.tfull[buf].done(seq // BUFS, 'epilogue/tfull'):
                    yield
            self.tempty[buf].arrive()                     # TMEM drained
            if pending is not None:                       # deferred publication, before this tile's stores
                s.publish(*pending)
                pending = None
            yield
            s.store(l, t)
            self.done_tiles += 1
            if l + 1 < s.L:
                pending = (t, l + 1)
            yield
        if pending is not None:
            s.publish(*pending)

    # ------------------------------------------------------------------ asynchronous hardware
    def retire_one(self):
        """the tensor pipe retires its oldest entry"""
        kind, *rest = self.pipe.pop(0)
        if kind == 'commit':
            rest[0].arrive()
            return
        seq, tap = rest
        l, k, t = self.tile(seq)
        if self.stage_tile[seq % STAGES] != seq:
            raise ProtocolError(f'cta {self.b}: MMA of seq {seq} reads stage holding {self.stage_tile[seq % STAGES]}')
        if self.slot[(9 * l + tap) % SLOTS] != (l, tap):
            raise ProtocolError(f'cta {self.b}: MMA layer {l} tap {tap} reads slot holding '
                                f'{self.slot[(9 * l + tap) % SLOTS]}')


class Sim:
    def __init__(self, tiles_x, tiles_y, n, G, L, seed, lookahead_blocks=False, gate_streamer=True, gate_early=True):
        self.tiles_x, self.tiles_y, self.n, self.L = tiles_x, tiles_y, n, L
        self.num_tiles = tiles_x * tiles_y * n
        self.G = min(G, self.num_tiles)
        self.lookahead_blocks, self.gate_streamer, self.gate_early = lookahead_blocks, gate_streamer, gate_early
        self.rng = random.Random(seed)
        self.flag = [0] * self.num_tiles                  # layers published per tile
        # SRNet buffer plan: layer 0: x(0) -> 1; odd layers: 1 -> 2; even layers > 0: 2 -> 1 (in place over
        # the residual).  version[buf][tile] = layer whose output the region holds (-1 = chain input / junk)
        self.src = [0] + [1 if l % 2 else 2 for l in range(1, L)]
        self.dst = [1] + [2 if l % 2 else 1 for l in range(1, L)]
        self.version = {0: [-1] * self.num_tiles, 1: [None] * self.num_tiles, 2: [None] * self.num_tiles}
        self.ctas = [Cta(self, b) for b in range(self.G)]

    def halo(self, t):
        per = self.tiles_x * self.tiles_y
        n, r = divmod(t, per)
        ty, tx = divmod(r, self.tiles_x)
        return [n * per + yy * self.tiles_x + xx
                for yy in range(max(ty - 1, 0), min(ty + 2, self.tiles_y))
                for xx in range(max(tx - 1, 0), min(tx + 2, self.tiles_x))]

    def check_halo_ready(self, l, t, when):
        want = l - 1
        for u in self.halo(t):
            have = self.version[self.src[l]][u]
            if have != want:
                raise ProtocolError(f'{when} of tile {t} layer {l}: halo tile {u} holds layer {have}, needs {want}')

    def store(self, l, t):
        self.version[self.dst[l]][t] = l

    def publish(self, t, layers_done):
        if self.flag[t] != layers_done - 1:
            raise ProtocolError(f'tile {t}: flag {self.flag[t]} -> {layers_done}')
        self.flag[t] = layers_done

    def run(self, max_steps=2_000_000):
        actors = []
        for c in self.ctas:
            actors += [c.producer(), c.mma(), c.streamer(), c.checker(), c.epilogue(0), c.epilogue(1)]
        live = list(actors)
        idle = 0
        for _ in range(max_steps):
            if not live and not any(c.pipe or c.loads for c in self.ctas):
                break
            before = self._snapshot()
            r = self.rng.random()
            hw = [c for c in self.ctas if c.pipe or c.loads]
            if hw and (r < 0.35 or not live):
                c = self.rng.choice(hw)
                if c.loads and (not c.pipe or self.rng.random() < 0.5):
                    c.loads.pop(self.rng.randrange(len(c.loads)))()      # loads complete in any order
                else:
                    c.retire_one()
            else:
                a = self.rng.choice(live)
                try:
                    next(a)
                except StopIteration:
                    live.remove(a)
            idle = idle + 1 if self._snapshot() == before else 0
            if idle > 20000:
                raise ProtocolError('deadlock: no progress in 20000 scheduling steps')
        else:
            raise ProtocolError('did not finish')
        assert sum(c.done_tiles for c in self.ctas) == self.num_tiles * self.L
        assert all(v == self.L - 1 for v in self.version[self.dst[self.L - 1]])

    def _snapshot(self):
        return (tuple(self.flag), tuple(c.done_tiles for c in self.ctas), tuple(len(c.pipe) for c in self.ctas),
                tuple(len(c.loads) for c in self.ctas), tuple(c.deps_ok for c in self.ctas),
                tuple(b.phase for c in self.ctas for b in c.full + c.tfull + c.wfull + c.wfree + [c.wstart, c.wearly]))


SHAPES = [
    # tiles_x, tiles_y, n, G, L
    (1, 1, 1, 4, 3),      # one tile, one CTA
    (3, 2, 1, 8, 4),      # one tile per CTA
    (3, 2, 1, 4, 5),      # 1-2 tiles per CTA
    (4, 3, 2, 5, 4),      # ~5 tiles per CTA, two images
    (4, 3, 1, 3, 6),      # 4 tiles per CTA
    (5, 2, 1, 1, 3),      # a single CTA owning everything
    (2, 2, 1, 3, 1),      # a chain of one layer
    (6, 2, 1, 2, 2),      # first tile of layer 1 is also the last-but-n of layer 0
]


@pytest.mark.parametrize('shape', SHAPES)
def test_chain_protocol_random_schedules(shape):
    for seed in range(16):
        Sim(*shape, seed=seed).run()


def test_model_catches_the_protocol_bugs():
    # (1) a BLOCKING look-ahead on the next tile's barriers before the current tile is committed
    #     deadlocks as soon as the next tile depends on the current one (one tile per CTA)
    with pytest.raises(ProtocolError, match='deadlock'):
        for seed in range(4):
            Sim(3, 2, 1, 8, 3, seed=seed, lookahead_blocks=True).run()
    # (2) without the wstart gate the streamer can complete a weight barrier's second phase before the
    #     MMA issuer has waited for the first one: a parity wait would then hang
    with pytest.raises(ProtocolError, match='phases past'):
        for seed in range(40):
            Sim(3, 2, 1, 4, 4, seed=seed, gate_streamer=False).run()
    # (3) without the wearly gate a streamer that is scheduled late (one tile per layer) finds the
    #     hand-over barrier of taps 4-8 completed twice
    with pytest.raises(ProtocolError, match='streamer/wfree'):
        for seed in range(40):
            Sim(1, 1, 1, 1, 4, seed=seed, gate_early=False).run()
