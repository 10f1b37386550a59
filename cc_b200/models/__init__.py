"""Model factory with the reference's names (models/__init__.py): ``getattr(models, name)(...)``
as train.py:245-255 does.  Only the architectures on the BASELINE.json path are built."""
from .DispResNet6 import DispResNet6
from .PoseNetB6 import PoseNetB6
from .MaskNet6 import MaskNet6
try:
    from .back2future import Model as Back2Future
except ImportError:       # pragma: no cover
    pass
