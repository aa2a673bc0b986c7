"""``LaneType`` -- the only live symbol of the reference's lane.py (planners/common/scenario/lane.py:8-12)."""
from enum import Enum


class LaneType(Enum):
    UNDEFINED = 0
    LEFT = 1
    EGO = 2
    RIGHT = 3
