"""Imported automatically by the interpreter when this directory is on PYTHONPATH: installs the demo's compatibility layer
(see dc_py2compat.py) before python/pose/pose_demo.py of the reference starts."""
import dc_py2compat

dc_py2compat.install()
