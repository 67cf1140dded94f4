"""fsb200 — B200-native backend for the Fengshen data-parallel pretraining step."""
__version__ = "0.1"
