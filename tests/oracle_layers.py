"""Oracle-side emulation of the reference's layer control flow (TEST INFRASTRUCTURE): the host
logic of QuantizeLayer.forward (qsparse/quantize.py:473-518) and PruneLayer.forward +
MagnitudePruningCallback.forward (qsparse/sparse.py:99-122,215-273) driving the oracle's
arithmetic.  Used to check every layer call of an end-to-end GPU run on identical inputs."""
import numpy as np

from oracle import oracle as orc


class OracleQuantize:
    """per-tensor Decimal / Scaler quantize layer (channelwise = -1)"""

    def __init__(self, bits, timeout, kind):
        self.bits, self.timeout, self.kind = bits, timeout, kind
        self.n = 0
        self.cb_t = 0
        self.weight = np.zeros(1, np.float32)
        self.quantized = False

    def forward(self, x, training=True):
        out = x
        if self.timeout > 0 and self.n >= self.timeout:
            if training:
                self.weight = orc.scale_ema(self.weight, orc.absmax(x, -1), self.bits, self.cb_t)
                self.cb_t += 1
                self.quantized = True
            if self.quantized:
                if self.kind == "decimal":
                    out = orc.fq_pow2_fwd(x, orc.scale_to_decimal(self.weight))
                else:
                    out = orc.fq_scaler_fwd(x, self.weight)
        if training and self.timeout > 0:
            self.n += 1
        return out


class OraclePrune:
    """structured magnitude pruning with running average; ``dimensions`` = the axes the mask keeps (any set, ref
    qsparse/sparse.py:236-241)"""

    def __init__(self, sparsity, start, interval, repetition, rampup=False, dimensions=(1,)):
        self.dimensions = set(dimensions)
        self.sparsity, self.start, self.interval, self.repetition = sparsity, start, interval, repetition
        self.schedules, self.rampup_interval = orc.prune_schedule(start, interval, repetition, rampup)
        self.n = 0
        self.cur = 0.0
        self.t = -1
        self.mask = None
        self.mag = None

    def forward(self, x):
        mshape = tuple(s if i in self.dimensions else 1 for i, s in enumerate(x.shape))
        if self.mask is None:
            self.mask = np.ones(mshape, bool)
        if self.n in self.schedules:
            self.cur = orc.ramp_sparsity(self.n, self.sparsity, self.start, self.interval, self.repetition,
                                         self.rampup_interval)
        if self.n >= self.start:
            if self.t == -1:
                self.mag = np.zeros(mshape, np.float32)
                self.t = 0
            self.mag = orc.magnitude_ema(self.mag, orc.squeeze_mean_abs(x, mshape), self.t)
            if orc.refresh_gate(self.t, self.cur, 1, float("inf"), True):
                self.mask, _ = orc.mask_given_importance(self.mag, self.cur)
            if self.dimensions == {1}:
                out = orc.mask_apply(x, self.mask.reshape(-1), 1)
            else:                       # x * mask by broadcasting (IEEE: x * 0.0 keeps the sign of x)
                out = x * self.mask.reshape(mshape).astype(np.float32)
            self.t += 1
        else:
            out = x
        self.n += 1
        return out
