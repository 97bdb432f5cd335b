from .prior_generator import PriorGenerator  # noqa: F401
from .prob_generator import ProbGenerator  # noqa: F401
