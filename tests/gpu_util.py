"""Helpers shared by the GPU parity tests: load an oracle State into the CUDA engine and compare."""
import numpy as np
import torch

from imm_b200 import _lib
from imm_b200.engine import IMMEngine
from imm_b200.utils.box import default_model_config
from oracle import imm_oracle as O


def rel_err(a, b):
  a = torch.as_tensor(a).detach().double().cpu()
  b = torch.as_tensor(b).detach().double().cpu()
  n = float(b.norm())
  return float((a - b).norm()) / (n if n > 0 else 1.0)


def make_pair(batch=2, n_maps=10, image_size=128, seed=0, precision=None, engine=_lib.ENGINE_AUTO,
              world_size=1):
  """Returns (engine on cuda:0, fp64 oracle state, fp32 oracle state, cpu inputs) with identical parameters."""
  st32 = O.init_state(O.State(n_maps=n_maps, image_size=image_size), seed=seed)
  st64 = st32.clone(torch.float64)
  eng = IMMEngine(default_model_config(n_maps), batch, image_size, 'cuda:0', precision=precision, engine=engine,
                  world_size=world_size)
  eng.load_state(st32.params, st32.buffers)
  eng.load_vgg_caffe_dict(O.synthetic_vgg_caffe_dict(seed + 1))
  inputs = O.synthetic_inputs(batch, image_size, seed=seed)
  return eng, st64, st32, inputs


def to_dev(inputs):
  return {k: v.cuda().contiguous() for k, v in inputs.items()}
