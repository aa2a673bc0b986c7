"""fiss_plus_planner_b200 -- B200-native Frenet trajectory sampling-and-scoring engine."""
__version__ = "0.1.0"
