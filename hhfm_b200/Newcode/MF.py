"""Drop-in for the model class of the reference's Newcode/MF.py (MF.py:43-149).  The reference file is a script
with a broken trainer (`args` global at :181, missing `evaluate` at :295); only the model API is kept."""
from hhfm_b200.models import MF  # noqa: F401
