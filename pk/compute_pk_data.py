"""Drop-in for the reference's pk/compute_pk_data.py: unwindowed P_l(k) of a catalogue.

    python pk/compute_pk_data.py <int> <paramfile>

Same arguments, .param format, file names and output layout as the reference script; the body runs
on the GPU through libsww.so (see spectra_without_windows_b200/drivers.py).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
from spectra_without_windows_b200.drivers import pk_data_main  # noqa: E402

pk_data_main(sys.argv)
