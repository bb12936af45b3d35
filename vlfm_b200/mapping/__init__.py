from .base_map import BaseMap  # noqa: F401
from .value_map import ValueMap, ValueMapBatch  # noqa: F401
from .obstacle_map import ObstacleMap  # noqa: F401
from .frontier_map import Frontier, FrontierMap  # noqa: F401
