"""``framefusion.utils`` of the reference (utils.py:9-57) -> ``framefusion_b200.utils``."""
from framefusion_b200.utils import TEXT_TOKEN, IGNORE_TOKEN, get_attr_by_name, scaled_dot_product_attention  # noqa: F401
