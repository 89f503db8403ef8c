"""``framefusion.main`` of the reference (main.py:8-380) -> ``framefusion_b200.main``."""
from framefusion_b200.main import FrameFusion, cosine_similarity, find_contigious_latter_index  # noqa: F401
from framefusion_b200.utils import TEXT_TOKEN, IGNORE_TOKEN  # noqa: F401
