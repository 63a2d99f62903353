"""Abstract model class -- host-side mirror of imm/models/base_model.py (BaseModel).

The reference builds TF graph nodes; here the same helpers keep their names and argument meaning but act
eagerly on the CUDA engine (imm_b200/engine.py).  file:line citations are under /root/reference."""


class BaseModel(object):
  num_instances = 0

  def __init__(self, dtype, name):
    self.dtype = dtype
    self._name = name
    self._avg_ops = []          # base_model.py:26-27 (moving-average ops)
    self._opts = None
    self._cost_avgs = {}        # name -> EMA value (tf.train.ExponentialMovingAverage(0.99), base_model.py:56-60)
    self._cost_raw = {}
    self._running = {}          # name -> value of the host-side _exp_running_avg helper
    self.__class__.num_instances += 1

  def _decay(self, scope=None):
    """Sum of the L2 weight-decay losses (base_model.py:33-37): computed on device by immb_total_loss."""
    raise NotImplementedError

  def _exp_running_avg(self, x, training_pl, init_val=0.0, rho=0.99, name='x'):
    """base_model.py:39-50: x_new = agg + (1 - rho) (x - agg); agg <- x_new when training_pl; returns x_new.  Host-side
    scalar version for callers of the helper; the perceptual-loss normalisers `<level>_agg` of the hot path live in the
    engine (engine.agg) and are advanced on device by immb_perceptual_finalize."""
    agg = self._running.get(name + '_agg', float(init_val))
    x_new = agg + (1.0 - rho) * (float(x) - agg)
    if training_pl:
      self._running[name + '_agg'] = x_new
    return x_new

  def _add_cost_summary(self, cost, name):
    """Raw + moving-average cost scalars (base_model.py:52-60); only for the first model instance."""
    if True:      # the reference only registers summaries for the first instance (num_instances == 1); EMAs are cheap here
      def update(cost=cost, name=name):
        v = float(cost() if callable(cost) else cost)
        self._cost_raw[name] = v
        prev = self._cost_avgs.get(name)
        # TF EMA with zero_debias=False initialises the shadow variable with the first value
        self._cost_avgs[name] = v if prev is None else prev - (1.0 - 0.99) * (prev - v)
        return self._cost_avgs[name]
      self._avg_ops.append(update)

  def _get_opts(self, training_pl):
    if self._opts is None:
      self._opts = {'dtype': self.dtype, 'wd': 1e-5, 'std': 0.01, 'training_pl': training_pl}   # base_model.py:62-69
    return self._opts

  def get_bnorm_ops(self, scope=None):
    """base_model.py:71-78 returns the grouped BN moving-average updates.  On the CUDA path they are applied by
    immb_bn_finalize inside the training forward pass, so the returned op is a no-op callable."""
    return lambda: None

  def conv_block(self, *args, **kwargs):
    """base_model.py:96-117 creates the variables of ONE conv block inside a TF graph.  Here every block of the path
    (shape, packed weight planes, BN scratch) is created by IMMEngine at construction from the reference's layer specs;
    a free-standing block has no buffers to run on -- documented as graph-only in INTEGRATION.md."""
    raise NotImplementedError('graph-only helper: conv blocks are instantiated by IMMEngine from the layer specs '
                              '(nn_utils.py:151-210 semantics); see IMMModel.encoder / simple_renderer for runnable sections')

  def build(self, inputs, training_pl):
    raise NotImplementedError
