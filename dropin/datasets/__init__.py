"""Import-path shim: `datasets.*` as the reference scripts spell it.  Only `datasets.audioset` is provided (a
synthetic AudioSet, see there); the other dataset modules keep importing from the reference checkout."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
