"""BASELINE config 1 on the GPU path: the drop-in single SimulatedNetworkEnv (one env per object, Python's own MT19937
stream handed to the device and back around every call), 400-step episodes with N(0,1) actions, steps/s incl. resets.
python tools/time_dropin.py [episodes]"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcc_rl_b200
from pcc_rl_b200 import network_sim
episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 5
random.seed(100)
env = network_sim.SimulatedNetworkEnv()
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
print("drop-in SimulatedNetworkEnv on the GPU: %d steps (%d episodes incl. resets) in %.2f s = %.0f steps/s "
      "(reference in Python: 1.3-1.5 k steps/s per core)" % (steps, episodes, dt, steps / dt))
