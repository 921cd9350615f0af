"""segma_b200 -- B200-native sliding-window frame-level inference path of arxaqapi/segma."""
__version__ = "0.1.0"
