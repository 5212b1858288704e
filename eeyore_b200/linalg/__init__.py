from .is_pos_def import is_pos_def
from ..stats.stats import nearest_pd  # noqa: F401
