"""TEST INFRASTRUCTURE ONLY — stand-in for the absent `matplotlib` (the reference's src/summary/diffusion_dcbase_summary.py:3,8
imports pyplot for one colour map used when dumping PNGs; src/main.py::test() imports that module)."""
