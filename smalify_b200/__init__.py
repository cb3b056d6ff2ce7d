"""B200-native SMAL fitting inner loop (drop-in for SMALify's SMALFitter path)."""
__version__ = "0.1.0"
