"""Worker for the column-sharded run (launched by torchrun from test_multigpu.py / by hand)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import clp_b200
from clp_b200 import generators as G
from clp_b200.sharding import broadcast_unique_id

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m, n, dens, seed = [float(x) for x in sys.argv[1:5]]
lp = G.random_sparse_lp(int(m), int(n), dens, int(seed))
s = clp_b200.ClpSimplex(); s.loadLP(lp)
uid = clp_b200.ClpSimplex.ncclUniqueId() if rank == 0 else np.zeros(128, dtype=np.uint8)
uid = broadcast_unique_id(uid, src=0)
s.setParameter("shardMinNnzPerRank", 0)  # small test problem: force the sharded path
s.setParameter("shardPanel", 1)          # ... including the row-sharded eta panel
s.initSharding(rank, world, uid)
if len(sys.argv) > 5:
    s.setMaximumIterations(int(sys.argv[5]))
st = s.dual()
out = {"rank": rank, "status": st, "objective": s.objectiveValue(), "iterations": s.numberIterations(),
       "known": lp.known_objective, "seconds": s.secondsInLoop()}
objs = [None] * world
dist.all_gather_object(objs, out)
if rank == 0:
    print("MULTIGPU_RESULT " + json.dumps(objs), flush=True)
dist.destroy_process_group()
