"""Multi-rank parity check of the training step (run under torchrun, one rank per GPU, NCCL):

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_step_check.py [per_rank_batch]

Every rank runs IMMEngine(world_size=N).train_step on ITS slice of one global batch with the real all-reduce of the
flat gradient bucket.  The LAST rank then runs oracle.train_step(n_towers=N) on the whole batch (train_multi,
cnn_train_multi.py:109-192: even split :132, tower-gradient mean :93-94, clip :95-98, apply :164, BN updates of the
last tower :155,166) and compares: averaged gradients, updated parameters, its own BN moving statistics / loss
normalisers (= the last tower's), its tower loss.  Prints DIST_STEP_CHECK_OK on success (exit code 0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
  from imm_b200.engine import IMMEngine
  from imm_b200.train import cnn_train_multi as tru
  from imm_b200.utils.box import default_model_config
  from oracle import imm_oracle as O
  from tests.gpu_util import rel_err
  per = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  rank, local_rank, world = tru.init_distributed('nccl')
  assert world > 1, 'run under torchrun with at least 2 ranks'
  torch.cuda.set_device(local_rank)
  K = 10
  st32 = O.init_state(O.State(n_maps=K, image_size=128), seed=0)
  eng = IMMEngine(default_model_config(K), per, 128, 'cuda:%d' % local_rank, world_size=world)
  eng.load_state(st32.params, st32.buffers)
  eng.load_vgg_caffe_dict(O.synthetic_vgg_caffe_dict(1))
  inputs = O.synthetic_inputs(per * world, 128, seed=0)
  mine = {k: v[rank * per:(rank + 1) * per].cuda().contiguous() for k, v in inputs.items()}
  eng.train_step(mine['image'], mine['future_image'], mine['mask'], allreduce=tru.average_gradients)
  torch.cuda.synchronize()
  # every rank must hold the same reduced gradient bucket and the same updated parameters
  ref_g, ref_p = eng.flat_g.clone(), eng.flat_p.clone()
  dist.broadcast(ref_g, src=0)
  dist.broadcast(ref_p, src=0)
  assert torch.equal(ref_g, eng.flat_g) and torch.equal(ref_p, eng.flat_p), 'replicas diverged'
  ok = True
  if rank == world - 1:
    st64 = st32.clone(torch.float64)
    r64 = O.train_step(st64, {k: v.double() for k, v in inputs.items()}, n_towers=world)
    tower_loss = float(r64['out']['loss']) + 0.0
    got = float(eng.total_loss.item())
    assert abs(got - tower_loss) / tower_loss < 1e-4, (got, tower_loss)
    worst = 0.0
    for k, g in r64['grads'].items():
      if float(g.abs().max()) < 1e-6:
        continue
      e = rel_err(eng.grads[k] / world, g)          # flat_g holds the tower SUM; the optimiser applies 1/N
      worst = max(worst, e)
      assert e < 2e-2, ('grad', k, e)
      big = g.abs() > 5e-2 * g.abs().max()
      upd_ref = (st64.params[k] - st32.params[k].double())[big]
      upd_gpu = (eng.params[k].cpu().double() - st32.params[k].double())[big]
      assert rel_err(upd_gpu, upd_ref) < 5e-2, ('update', k)
    for k, v in st64.buffers.items():
      assert rel_err(eng.buffers[k], v) < 1e-4, ('buffer', k)
    print('dist_step_check: %d ranks x %d pairs, worst gradient rel. error %.2e' % (world, per, worst))
  dist.barrier()
  if rank == world - 1 and ok:
    print('DIST_STEP_CHECK_OK')
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
