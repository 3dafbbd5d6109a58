"""pyranges_b200 -- B200-native engine for the pyranges interval hot path."""
__version__ = "0.1.0"
