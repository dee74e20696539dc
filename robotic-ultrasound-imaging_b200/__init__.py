"""B200-native batched Ultrasound environment (hot path of
hermanjakobsen/robotic-ultrasound-imaging).

The directory name carries a hyphen; import it as ``rui_b200`` (alias package
at the repo root) or through ``importlib.import_module``.
"""
__version__ = "0.1.0"
