"""The model-class boundary on the GPU: sub-builders, `tensors` dict, piecewise model == build()."""
import pytest
import torch

from imm_b200.models.imm_model import IMMModel
from imm_b200.utils.box import default_model_config
from imm_b200.utils.synthetic import synthetic_inputs, synthetic_vgg_caffe_dict
from oracle import imm_oracle as O
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu


def test_sub_builders_reproduce_the_full_forward():
  """image_encoder / pose_encoder / simple_renderer (imm_model.py:220,233,154) run on their own and chained by hand give
  what build() gives; `tensors` carries the reference's keys incl. the colourised `pose_embedding` (:463-465,480-485)."""
  B = 2
  m = IMMModel(default_model_config(10), vgg_data=synthetic_vgg_caffe_dict(1), seed=11)
  inp = {k: v.cuda() for k, v in synthetic_inputs(B, 128, seed=4).items()}
  _, loss, avg_ops, tensors = m.build(inp, False, output_tensors=True)
  assert set(tensors) >= {'image', 'future_image', 'mask', 'future_im', 'im', 'pose_embedding', 'future_im_pred', 'gauss_yx'}
  assert tuple(tensors['pose_embedding'].shape) == (B, 128, 128, 3) and float(tensors['pose_embedding'].max()) <= 1.0
  pred_full, mu_full = tensors['future_im_pred'].clone(), tensors['gauss_yx'].clone()
  loss_full = float(loss.item())
  # against the oracle's forward in inference mode
  st = O.init_state(O.State(n_maps=10), seed=0)
  # piecewise
  feats = m.image_encoder(inp['image'], False)
  assert [tuple(f.shape[1:]) for f in feats] == [(128, 128, 3), (128, 128, 32), (64, 64, 64), (32, 32, 128), (16, 16, 256)]
  mu, maps = m.pose_encoder(inp['future_image'], False, n_maps=10, gauss_mode='rot', map_sizes=[16, 128])
  assert torch.equal(mu, mu_full) and [tuple(g.shape) for g in maps] == [(B, 16, 16, 10), (B, 128, 128, 10)]
  assert set(m.get_collection('tensors')) == {'heatmaps', 'gauss_y_prob', 'gauss_x_prob'}
  joint = torch.cat([feats[-1], maps[0]], dim=-1)
  pred = m.simple_renderer({16: joint}, False, n_final_out=3, final_res=128)
  assert rel_err(pred, pred_full) < 1e-5
  # model() + loss() == build()
  m._get_opts(False)
  pred2, mu2, emb = m.model(inp['image'], inp['future_image'])
  assert torch.equal(mu2, mu_full) and rel_err(pred2, pred_full) < 1e-6 and len(emb) == 4
  loss2 = m.loss(pred2, inp['future_image'], mu2, emb, 'costs', False, loss_mask=inp['mask'])
  assert abs(float(loss2.item()) - loss_full) <= 1e-6 * abs(loss_full)
  assert abs(float(m._decay().item()) - float(m.engine.weights_loss.item())) == 0.0
  assert m._exp_running_avg(2.0, True, init_val=1.0, name='t') == pytest.approx(1.01)
  assert m._exp_running_avg(2.0, False, init_val=1.0, name='t') == pytest.approx(1.01 + 0.01 * (2.0 - 1.01))
