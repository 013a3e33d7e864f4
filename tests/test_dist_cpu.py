"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: the num_boxes all-reduce of SetCriterion
 and the rank behaviour of bench.py's reference arm."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, counts, out):
    sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detr_b200 import HungarianMatcher, SetCriterion, pack_targets
    labels = [torch.zeros(c, dtype=torch.int64) for c in counts[rank]]
    boxes = [torch.rand(c, 4) for c in counts[rank]]
    pt = pack_targets(labels, boxes, 100, torch.device("cpu"))
    crit = SetCriterion(5, HungarianMatcher(1, 5, 2))
    nb = crit._num_boxes(pt, torch.device("cpu"))
    off = SetCriterion(5, HungarianMatcher(1, 5, 2), sync_num_boxes=False)._num_boxes(pt, torch.device("cpu"))
    out[rank] = (float(nb), off is None, pt.gt_off.tolist(), pt.match_off.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("counts,expect", [(([3, 0, 5], [10, 2, 0]), 10.0), (([0], [0]), 1.0), (([150], [1]), 75.5)])
def test_num_boxes_allreduce_gloo(counts, expect):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, counts, out), nprocs=2, join=True)
    assert out[0][0] == out[1][0] == pytest.approx(expect)      # mean over ranks of the local sums, clamped at 1
    assert out[0][1] and out[1][1]                              # disabled -> kernel falls back to the local count
    assert out[0][2] == [0] + list(torch.tensor(counts[0]).cumsum(0).tolist())
    assert out[0][3] == [0] + list(torch.tensor([min(c, 100) for c in counts[0]]).cumsum(0).tolist())


def test_single_process_has_no_allreduce():
    sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
    from detr_b200 import HungarianMatcher, SetCriterion, pack_targets
    pt = pack_targets([torch.zeros(2, dtype=torch.int64)], [torch.rand(2, 4)], 100, torch.device("cpu"))
    assert SetCriterion(5, HungarianMatcher(1, 5, 2))._num_boxes(pt, torch.device("cpu")) is None


def test_reference_arm_nonzero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_pack_targets_validation():
    sys.path.insert(0, os.path.join(ROOT, "detr-object-detection_b200"))
    from detr_b200 import pack_targets
    with pytest.raises(ValueError):
        pack_targets([torch.zeros(2, dtype=torch.int64)], [torch.rand(3, 4)], 10, torch.device("cpu"))
    pt = pack_targets([], [], 10, torch.device("cpu"))
    assert pt.batch == 0 and pt.total == 0 and pt.max_count == 0
