"""Stand-in for torchdiffeq (out of scope; only the names are needed to import flowcon.CNF)."""


def odeint(*args, **kwargs):
    raise NotImplementedError("torchdiffeq is stubbed")


def odeint_adjoint(*args, **kwargs):
    raise NotImplementedError("torchdiffeq is stubbed")
