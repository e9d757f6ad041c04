"""Benchmark denominators (NOT product code, never imported by spectral_connectivity_b200):

* ``ref_arm``  -- drives the UNMODIFIED reference package from ``baseline/_ref`` (git-ignored install made by
  ``__graft_entry__.build()`` in the build container; it travels to the GPU box with the snapshot) through the
  reference's own public API on the host cores;
* ``replay``   -- a torch replay of the library-call sequence the reference's CuPy backend would issue
  (cuFFT / cuBLAS / cuSOLVER underneath), the stand-in for the CuPy denominator (cupy is not installable here).
"""
