"""GPU, world size 2 over NCCL (skipped on a single-GPU box; run with `gpurun --gpus 2`): RESULTS of the batch-sharded
step, not only its speed --
  * the all-reduced flat gradient buffer x 1/world equals the mean of the two ranks' local gradients (gathered)
  * after eager steps AND after graph-replayed steps (the NCCL all-reduce and the multi-tensor Adam are recorded inside
    the step's CUDA graph) every rank holds bit-identical parameters
  * the replayed run's losses are finite and the optimizer step counters advanced identically
The reference has no distributed code; this is the data-parallel form of train...triplet.py:171-237."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import weights

pytestmark = pytest.mark.gpu
CFG_I = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
CFG_S = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    trainer = None
    try:
        import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
        from cooperative_training_and_latent_space_data_augmentation_b200 import training
        assert pkg.conv_blocks.get_precision() == "kernel"
        torch.manual_seed(100 + rank)                       # different initial weights per rank: the broadcast must fix it
        solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
        img, lab, _ = weights.synthetic_batch(8, 64, 64, seed=2)
        trainer = pkg.GraphedCooperativeTrainer(solver, 8, seed=3, image_cfg=CFG_I, seg_cfg=CFG_S, eager_steps=2)
        img, lab = trainer.local_slice(img).cuda(), trainer.local_slice(lab).cuda()
        res = {"sync0": trainer.params_in_sync()}
        # (1) two eager + four replayed steps
        losses = []
        for _ in range(6):
            losses.append(float(trainer.step(img, lab)['loss']))
        torch.cuda.synchronize()
        res.update(sync=trainer.params_in_sync(), losses=losses, captured=len(trainer.captured),
                   steps=solver.flat_adam.steps.tolist(), capture_collective=trainer.capture_collective)
        # (2) gradient exchange of one eager step without the optimizer (on the trainer's stream: collectives recorded
        # into a graph and eager ones must not be mixed across streams)
        with torch.cuda.stream(trainer.stream):
            training.cooperative_step(solver, img, lab, CFG_I, CFG_S, optimize=False)
            local = solver.flat_adam.flat_grads.clone()
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            want = torch.stack(gathered).double().mean(0)
            trainer.bucket.all_reduce_sum()
            got = trainer.bucket.mean_gradients().double()
            res["grad_err"] = float((got - want).abs().max() / want.abs().max())
            res["grad_differs_between_ranks"] = float((gathered[0] - gathered[1]).abs().max()) > 0
        out[rank] = res
    finally:
        if trainer is not None:
            trainer.close()                                 # graphs that recorded NCCL kernels go before the communicator
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_step_keeps_replicas_identical():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        res = out[r]
        assert res["sync0"] and res["sync"], res
        assert res["grad_err"] < 1e-6 and res["grad_differs_between_ranks"], res
        assert res["captured"] >= 1 and res["capture_collective"], res
        assert all(l == l and abs(l) < 1e4 for l in res["losses"]), res
        assert res["steps"] == [6.0] * 5, res
    assert out[0]["losses"] != out[1]["losses"]            # every rank logs the loss of its OWN shard
