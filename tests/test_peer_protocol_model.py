"""Executable model of the peer-memory exchange protocol (rigid_body_light_b200/csrc/rbl_peer.cuh, rbl_comm.h):
`world` threads play the ranks; every buffer a real rank would map from its peers is a shared Python object that
carries the product number it was written for; epochs are plain integers raised after the data, polled by the
waiting side.  The model executes exactly the sequence of `prod_M` in peer mode --

    push my lambda slice into EVERY rank's lambda region -> raise "lambda ready" everywhere, wait for everyone's
    -> product (consumes my lambda region, OVERWRITES my partial product) -> raise "partial ready", wait
    -> pull my rows of everyone's partial product

-- with random and adversarial delays, and checks the two hazards the header argues cannot happen without an extra
barrier: a lambda slot overwritten before its owner's product consumed it, and a partial product overwritten before
every rank's reduce has pulled it.  A mutated protocol (one wait removed) must trip the checker: the negative
controls show the model can see the hazard at all.  CPU only; tests the protocol, not the CUDA code."""
import random
import threading
import time

import pytest


class Hazard(AssertionError):
    pass


class Rank:
    def __init__(self, world):
        self.lam = [0] * world           # lam[src]: product number the slice of rank `src` was pushed for
        self.lam_consumed = 0            # last product whose lambda region this rank's product kernel has read
        self.mbuf = 0                    # product number of this rank's partial product
        self.mbuf_reads = {}             # product number -> how many ranks have pulled it
        self.epoch = [[0] * world, [0] * world]  # [kind][src]: 0 = lambda ready, 1 = partial ready
        self.lock = threading.Lock()


def run_model(world, products, skip_lambda_wait=False, skip_partial_wait=False, slow_rank=None, seed=0):
    ranks = [Rank(world) for _ in range(world)]
    errors = []
    deadline = time.monotonic() + 20.0

    def wait(me, kind, k):
        while any(ranks[me].epoch[kind][r] < k for r in range(world)):
            if errors or time.monotonic() > deadline:
                raise Hazard("stopped")
            time.sleep(0)

    def body(me):
        rng = random.Random(seed * 100 + me)

        def dally(scale=1.0):
            t = rng.random() * 2e-4 * scale
            if slow_rank == me:
                t += 1e-3
            time.sleep(t)

        try:
            for k in range(1, products + 1):
                # 1. push (stores into every rank's lambda region, own included)
                for r in range(world):
                    with ranks[r].lock:
                        if ranks[r].lam_consumed < k - 1:
                            raise Hazard(f"rank {me} overwrites lambda slot {me} of rank {r} for product {k} before rank {r} "
                                         f"consumed product {k - 1} (it has consumed {ranks[r].lam_consumed})")
                        ranks[r].lam[me] = k
                dally()
                # 2. hand-shake "lambda ready"
                for r in range(world):
                    ranks[r].epoch[0][me] = k
                if not skip_lambda_wait:
                    wait(me, 0, k)
                dally()
                # 3. product: consumes the lambda region, overwrites the partial product
                with ranks[me].lock:
                    if not skip_lambda_wait and any(v != k for v in ranks[me].lam):
                        raise Hazard(f"rank {me} product {k} reads lambda slots {ranks[me].lam}")
                    ranks[me].lam_consumed = k
                    if k > 1 and ranks[me].mbuf_reads.get(k - 1, 0) < world:
                        raise Hazard(f"rank {me} overwrites its partial product {k - 1} after only "
                                     f"{ranks[me].mbuf_reads.get(k - 1, 0)} of {world} ranks pulled it")
                    ranks[me].mbuf = k
                dally(3.0)
                # 4. hand-shake "partial ready"
                for r in range(world):
                    ranks[r].epoch[1][me] = k
                if not skip_partial_wait:
                    wait(me, 1, k)
                dally()
                # 5. pull my rows of everyone's partial product
                for r in range(world):
                    with ranks[r].lock:
                        if not skip_partial_wait and ranks[r].mbuf != k:
                            raise Hazard(f"rank {me} reduce {k} pulls partial product {ranks[r].mbuf} of rank {r}")
                        ranks[r].mbuf_reads[ranks[r].mbuf] = ranks[r].mbuf_reads.get(ranks[r].mbuf, 0) + 1
                dally()
        except Hazard as exc:
            errors.append(str(exc))

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return [e for e in errors if e != "stopped"]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_protocol_has_no_reuse_hazard(world):
    for seed, slow in ((0, None), (1, 0), (2, world - 1)):
        assert run_model(world, 60, slow_rank=slow, seed=seed) == []


def test_negative_control_without_the_lambda_wait():
    """Without the wait of the 'lambda ready' hand-shake a fast rank overwrites its partial product before the slow
    rank has pulled it (the wait is what orders the overwrite after every peer's previous reduce)."""
    found = []
    for seed in range(6):
        found += run_model(3, 40, skip_lambda_wait=True, slow_rank=1, seed=seed)
        if found:
            break
    assert found, "the model did not detect the hazard of the mutated protocol"


def test_negative_control_without_the_partial_wait():
    """Without the wait of the 'partial ready' hand-shake a fast rank runs ahead into the next push while a slow
    rank has not consumed its lambda region (or pulls a partial product that is not there yet)."""
    found = []
    for seed in range(6):
        found += run_model(3, 40, skip_partial_wait=True, slow_rank=2, seed=seed)
        if found:
            break
    assert found, "the model did not detect the hazard of the mutated protocol"
