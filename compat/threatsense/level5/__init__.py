"""compat/ shadows single modules of this reference package; the rest of it stays importable from the reference's src/."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
