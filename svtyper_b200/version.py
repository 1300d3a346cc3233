"""Version string of the drop-in (the reference it mirrors is hall-lab/svtyper v0.7.1, svtyper/version.py:2)."""
__version__ = "0.7.1+b200.2"
__author__ = "svtyper_b200 contributors"
