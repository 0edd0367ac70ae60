"""Link-parameter sampling on the host, with the reference's formulas.

create_new_links_and_senders (gym/network_sim.py:454-467) draws, in this order,
    bw    = U(min_bw, max_bw)                 packets/s
    lat   = U(min_lat, max_lat)               s
    queue = 1 + int(exp(U(min_queue, max_queue)))   packets
    loss  = U(min_loss, max_loss)
    rate0 = U(0.3, 1.5) * bw
They are INPUTS of the device path.  For the batched env they are a pure function
of (seed, episode index, GLOBAL env id), so that an env batch sharded over several GPUs gets the
same parameters as the same batch on one GPU.
"""
import numpy as np


class LinkRanges(object):
    """Parameter ranges; defaults are the reference's (network_sim.py:355-358)."""

    def __init__(self, bw=(100.0, 500.0), lat=(0.05, 0.5), queue=(0.0, 8.0), loss=(0.0, 0.05),
                 start_factor=(0.3, 1.5)):
        self.bw, self.lat, self.queue, self.loss, self.start_factor = bw, lat, queue, loss, start_factor

    def max_queue_packets(self):
        return 1 + int(np.exp(self.queue[1]))


def sample_link_params(seed, episode, global_ids, n_global, ranges=None):
    """Returns dict of float64/int64 arrays (bw, lat, queue, loss, start_rate) for the envs with the
    given global ids in episode `episode`."""
    r = ranges or LinkRanges()
    g = np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(episode)]))
    u = g.random((int(n_global), 5))[np.asarray(global_ids, dtype=np.int64)]
    bw = r.bw[0] + (r.bw[1] - r.bw[0]) * u[:, 0]
    lat = r.lat[0] + (r.lat[1] - r.lat[0]) * u[:, 1]
    queue = 1 + np.exp(r.queue[0] + (r.queue[1] - r.queue[0]) * u[:, 2]).astype(np.int64)
    loss = r.loss[0] + (r.loss[1] - r.loss[0]) * u[:, 3]
    start_rate = (r.start_factor[0] + (r.start_factor[1] - r.start_factor[0]) * u[:, 4]) * bw
    return dict(bw=bw, lat=lat, queue=queue, loss=loss, start_rate=start_rate)


def validate_link_params(bw, lat, queue, loss, rates):
    """Explicit link parameters handed to reset(): the device loops assume a positive finite bandwidth and sending
    rate (a pacing timer that does not advance would never reach the end of the MI), a non-negative delay and queue,
    and a loss probability.  The reference never checks -- its own sampler cannot produce anything else -- but a
    batched env should refuse garbage instead of occupying a GPU.  Raises ValueError naming the offending field."""
    def bad(name, a, ok):
        a = np.asarray(a)
        m = ~ok(a)
        if m.any():
            i = int(np.flatnonzero(m.reshape(-1))[0])
            raise ValueError("link parameter %s[%d] = %r is out of range" % (name, i, a.reshape(-1)[i].item()))
    bad("bw", bw, lambda a: np.isfinite(a) & (a > 0))
    bad("lat", lat, lambda a: np.isfinite(a) & (a >= 0))
    bad("queue", queue, lambda a: np.asarray(a) >= 0)
    bad("loss", loss, lambda a: (a >= 0) & (a <= 1))
    bad("start_rate", rates, lambda a: np.isfinite(a) & (a > 0))
