"""``framefusion.interface`` of the reference (interface.py:47-214) -> ``framefusion_b200.interface``."""
from framefusion_b200.interface import apply_framefusion, get_token_type, replace_framefusion_forward  # noqa: F401
from framefusion_b200.main import FrameFusion  # noqa: F401
from framefusion_b200.utils import TEXT_TOKEN, IGNORE_TOKEN, get_attr_by_name  # noqa: F401
