from .vae import VAESolver
from .intro import IntroSolver
from .tc import TCSovler, TCLossMixin
from .intro_tc import IntroTCSovler

__all__ = ["VAESolver", "IntroSolver", "TCSovler", "IntroTCSovler", "TCLossMixin"]
