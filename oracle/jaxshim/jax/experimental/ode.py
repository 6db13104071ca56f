"""``jax.experimental.ode`` is only used by CompiledRungeKutta, which is out of scope."""


def runge_kutta_step(*args, **kwargs):
    raise NotImplementedError("jax.experimental.ode is not part of the shim")
