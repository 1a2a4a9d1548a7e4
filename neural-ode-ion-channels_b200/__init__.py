"""B200-native batched ``odeint`` for the hERG/IKr neural-ODE models (NN-f / NN-d)."""
from . import protocols  # noqa: F401
