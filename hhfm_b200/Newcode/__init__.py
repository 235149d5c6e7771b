"""Drop-in replacement for the reference's `Newcode/` directory: same module names, entry points and classes
(`FM_main`, `M7_main`, `BPR_main`, ... / `FM`, `OUR`, `BPR`, `MF`, `LoadData`), running on libhhfm_sm100.so.

Both import styles of the reference work: `import Newcode.NewLoadData` (put `hhfm_b200/` on sys.path, FM.py:15)
and the flat `from FM import FM_main` of main.py:9 (put `hhfm_b200/Newcode/` on sys.path).
"""
