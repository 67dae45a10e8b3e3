"""dev tool: two (or more) 10M-point tiles of the bench scene in ONE process on one GPU through the mailbox path
(nimrud_b200.distributed.process_tiles_local); meant to be run under ncu to see every kernel of a tile step."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import distributed as nd, synth
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gather = len(sys.argv) > 3 and sys.argv[3] == "gather"
extent = math.sqrt(n / 40.0)
tiles = [synth.urban_scene(n, seed=20 + r, device="cuda", origin=((r % 2) * extent, (r // 2) * extent)) for r in range(world)]
mbs = nd.HaloMailbox.local_set(world, ["cuda"] * world, torch.float32, n)
for it in range(3):
    outs = nd.process_tiles_local(tiles, EDGES, RADII, mailboxes=mbs, gather=gather)
    torch.cuda.synchronize()
print("ok", [float(o[:, 0].mean()) for o in outs])
