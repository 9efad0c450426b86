"""Model check of the hand-over protocol of the warp-specialised loss + gradient kernel (csrc/fusion_loss_ws.cuh), on the CPU.

The kernel's four warp groups (G0: TMA ring + V, G1: H, G2: S, G3: B1 + B2) exchange tiles through shared memory with
hardware named barriers (bar.arrive by the producer group / bar.sync by the consumer group) and, for the input ring, mbarriers.
This test restates each group's per-batch sequence of waits / arrivals / buffer accesses exactly as the kernel issues them and
runs the four programs under thousands of random interleavings, asserting what the hardware needs and what correctness needs:

  * no deadlock for any batch count (1 .. 14) and any first emitting batch;
  * every named barrier has at most ONE phase in flight (a second bar.arrive before the consumers' bar.sync would be counted
    into the same phase and release the barrier without them);
  * every read sees the tile it expects (vbuf / cbuf / gbuf / tbuf / ring slot of the right batch) and no buffer is overwritten
    while a reader of its previous content is still due.

It checks the PROTOCOL (the order of the lines marked in the kernel), not the CUDA code: when the kernel's hand-over changes,
this model has to change with it.  The GPU-side evidence is tests/test_single_pass_gpu.py (bit-comparisons on a shape sweep)
and compute-sanitizer racecheck / synccheck (profiles/r2b_sanitizer.txt).
"""
import random

import pytest

SLOTS = 7            # kWsSlots: ring slots of 8 input rows


class Deadlock(Exception):
    pass


class Sim:
    def __init__(self, nb, first_emit, seed):
        self.nb, self.first_emit = nb, first_emit
        self.rng = random.Random(seed)
        self.named = {}                 # named barrier id -> [producer arrivals, consumer syncs]
        self.full = set()               # ring groups whose load has completed (ring_full phases)
        self.released = {}              # ring group -> set of groups ('G2', 'G3') that arrived on ring_empty
        self.buf = {}                   # buffer instance -> (content tag, set of pending readers)
        self.tbuf_sync = [0, 0]         # G3-internal bar.sync: trivially satisfied in a one-thread-per-group model

    # ---- primitive actions: return False when blocked
    def arrive(self, bar):
        st = self.named.setdefault(bar, [0, 0])
        assert st[0] - st[1] == 0, f'second arrival on named barrier {bar} before its consumers synced'
        st[0] += 1
        return True

    def sync(self, bar):
        st = self.named.setdefault(bar, [0, 0])
        if st[0] == st[1]:
            return False
        st[1] += 1
        return True

    def write(self, name, tag, readers):
        if name in self.buf:
            old_tag, pending = self.buf[name]
            assert not pending, f'{name}: overwritten with {tag} while {pending} still have to read {old_tag}'
        self.buf[name] = (tag, set(readers))
        return True

    def read(self, name, tag, who):
        assert name in self.buf, f'{who} reads {name} before anything was written'
        cur, pending = self.buf[name]
        assert cur == tag, f'{who} expects {tag} in {name}, finds {cur}'
        pending.discard(who)
        return True

    # ---- the four programs (generators of thunks; a thunk returns False while blocked)
    def emit(self, b):
        return b >= self.first_emit

    def ring_readers(self, g):
        # who still has to read ring group g: V(b) reads groups b..b+2, S(b) reads b..b+2, the combine of B2(b) reads group b
        r = set()
        for b in range(max(0, g - 2), min(self.nb, g + 1)):
            r.add(('G0', b))
            r.add(('G2', b))
        if g < self.nb and self.emit(g):
            r.add(('G3', g))
        return r

    def issue(self, g):
        def wait_slot():
            if g < SLOTS:
                return True
            return self.released.get(g - SLOTS, set()) >= {'G2', 'G3'}
        yield wait_slot
        yield lambda: self.write(('ring', g % SLOTS), g, self.ring_readers(g))
        yield lambda: (self.full.add(g), True)[1]

    def g0(self):
        nb = self.nb
        for g in (0, 1, 2):
            yield from self.issue(g)
        yield lambda: 0 in self.full
        yield lambda: 1 in self.full
        for b in range(nb):
            if b + 1 < nb:
                yield from self.issue(b + 3)
            yield lambda b=b: (b + 2) in self.full
            if b >= 2:
                yield lambda b=b: self.sync(('VEMPTY', b & 1))
            for g in (b, b + 1, b + 2):
                yield lambda b=b, g=g: self.read(('ring', g % SLOTS), g, ('G0', b))
            yield lambda b=b: self.write(('vbuf', b & 1), b, {'G1'})
            yield lambda b=b: self.arrive(('VFULL', b & 1))

    def g1(self):
        nb = self.nb
        for b in range(nb):
            yield lambda b=b: self.sync(('VFULL', b & 1))
            if b >= 2:
                yield lambda b=b: self.sync(('CEMPTY', b & 1))
            yield lambda b=b: self.read(('vbuf', b & 1), b, 'G1')
            yield lambda b=b: self.write(('cbuf', b & 1), b, {'G3'})
            if b + 2 < nb:
                yield lambda b=b: self.arrive(('VEMPTY', b & 1))
            yield lambda b=b: self.arrive(('CFULL', b & 1))

    def g2(self):
        nb = self.nb
        yield lambda: 0 in self.full
        yield lambda: 1 in self.full
        for b in range(nb):
            yield lambda b=b: (b + 2) in self.full
            if b >= 2:
                yield lambda b=b: self.sync(('GEMPTY', b & 1))
            for g in (b, b + 1, b + 2):
                yield lambda b=b, g=g: self.read(('ring', g % SLOTS), g, ('G2', b))
            yield lambda b=b: self.write(('gbuf', b & 1), b, {'G3'} if self.emit(b) else set())
            yield lambda b=b: self.arrive(('GFULL', b & 1))
            yield lambda b=b: (self.released.setdefault(b, set()).add('G2'), True)[1]

    def g3(self):
        nb = self.nb
        for b in range(nb):
            yield lambda b=b: self.sync(('CFULL', b & 1))
            yield lambda b=b: self.read(('cbuf', b & 1), b, 'G3')
            if b + 2 < nb:
                yield lambda b=b: self.arrive(('CEMPTY', b & 1))
            if self.emit(b):
                yield lambda b=b: self.write('tbuf', b, {'G3'})
            yield lambda b=b: self.sync(('GFULL', b & 1))
            if self.emit(b):
                yield lambda b=b: b in self.full
                yield lambda b=b: self.read('tbuf', b, 'G3')
                yield lambda b=b: self.read(('ring', b % SLOTS), b, ('G3', b))
                yield lambda b=b: self.read(('gbuf', b & 1), b, 'G3')
            if b + 2 < nb:
                yield lambda b=b: self.arrive(('GEMPTY', b & 1))
            yield lambda b=b: (self.released.setdefault(b, set()).add('G3'), True)[1]

    def run(self):
        progs = {'G0': self.g0(), 'G1': self.g1(), 'G2': self.g2(), 'G3': self.g3()}
        pending = {k: next(p, None) for k, p in progs.items()}
        steps = 0
        while any(v is not None for v in pending.values()):
            live = [k for k, v in pending.items() if v is not None]
            self.rng.shuffle(live)
            for k in live:
                if pending[k]():
                    pending[k] = next(progs[k], None)
                    # let a group run ahead for a random stretch: long leads are where hand-over bugs hide
                    for _ in range(self.rng.randint(0, 12)):
                        if pending[k] is None or not pending[k]():
                            break
                        pending[k] = next(progs[k], None)
                    break
            else:
                raise Deadlock(f'nb={self.nb} first_emit={self.first_emit}: all of {live} are blocked')
            steps += 1
            assert steps < 100000
        # the gbuf of a non-emitting batch is written without a reader; everything else has been consumed
        for name, (tag, pend) in self.buf.items():
            assert not pend or name[0] == 'ring', (name, tag, pend)


@pytest.mark.parametrize('nb', list(range(1, 15)))
def test_hand_over_protocol_has_no_deadlock_no_double_phase_and_no_stale_tile(nb):
    for first_emit in range(0, min(nb, 3)):
        for seed in range(120):
            Sim(nb, first_emit, seed * 7919 + nb).run()


def test_the_model_catches_a_missing_empty_barrier():
    """Sanity of the checker itself: dropping the vbuf `empty` hand-over must be detected (an overwrite before H has read, or a
    second VFULL arrival in the same phase)."""
    class Broken(Sim):
        def sync(self, bar):
            if bar[0] == 'VEMPTY':
                return True                      # G0 does not wait for H(b - 2)
            return super().sync(bar)

        def arrive(self, bar):
            if bar[0] == 'VEMPTY':
                return True
            return super().arrive(bar)

    caught = 0
    for seed in range(200):
        try:
            Broken(8, 1, seed).run()
        except AssertionError:
            caught += 1
    assert caught > 0
