"""B200-native mirrors of ``pb_sed.models`` (weak_label.CRNN = FBCRNN, strong_label.CRNN = BiCRNN)."""
from . import base, weak_label, strong_label  # noqa: F401
