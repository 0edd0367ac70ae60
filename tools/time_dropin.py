"""BASELINE config 1 on the GPU path: the drop-in single SimulatedNetworkEnv (one env per object, CPython's MT19937
stream on the device, exchanged with Python's `random` at every reset), 400-step episodes with N(0,1) actions, steps/s
incl. resets.   python tools/time_dropin.py [episodes] [strict]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_rl_b200
from pcc_rl_b200 import network_sim
episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 5
random.seed(100)
strict = len(sys.argv) > 2 and sys.argv[2] == "strict"
env = network_sim.SimulatedNetworkEnv(strict_rng=strict)
arng = random.Random(101)
env.reset()
for _ in range(20):
    env.step([arng.gauss(0, 1)])
t0 = time.perf_counter()
steps = 0
for ep in range(episodes):
    env.reset()
    done = False
    while not done:
        obs, r, done, _ = env.step([arng.gauss(0, 1)])
        steps += 1
dt = time.perf_counter() - t0
print("drop-in SimulatedNetworkEnv on the GPU (%s RNG exchange): %d steps (%d episodes incl. resets) in %.2f s = %.0f steps/s"
      % ("per-step" if strict else "per-reset", steps, episodes, dt, steps / dt))
try:   # the unmodified Python reference on one core of this box, same episode pattern
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_timing
    r = ref_timing.time_reference(1, 100, 2000)
    if r:
        print("unmodified Python reference on one core of this box: %.0f steps/s" % r["value"])
except Exception as ex:
    print("reference not timed:", ex)
