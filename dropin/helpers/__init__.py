"""Import-path shim: `helpers.*` as the reference scripts spell it -> efficientat_b200.
Sub-modules this repository does not replace (models.ensemble, models.mn.attention_pooling, ...) fall through
to the reference checkout, which must come later on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
