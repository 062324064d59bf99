"""Backward passes of the autograd Functions (K3).  Filled in with the
backward kernels; until then training raises instead of silently falling
back to PyTorch autograd."""


def egnn_layer_backward(ctx, d_h, d_x, d_m):
    raise NotImplementedError('pvs_egnn_layer_bwd is not built yet')


def linear_backward(ctx, d_out):
    raise NotImplementedError('pvs_linear_bwd is not built yet')


def mean_pool_backward(ctx, d_pooled):
    raise NotImplementedError('pvs_mean_pool_bwd is not built yet')
