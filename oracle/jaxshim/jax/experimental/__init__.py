from . import jet, ode  # noqa: F401
