"""``jax.ops.index_update`` was removed from JAX long ago; only the (broken) wave_2d problem references it."""


def index_update(*args, **kwargs):
    raise NotImplementedError("jax.ops.index_update does not exist in current JAX either")
