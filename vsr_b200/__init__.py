"""Import shim: the product package lives in `video-subtitle-remover_b200/` (not a valid Python
identifier), this makes it importable as `vsr_b200`."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "video-subtitle-remover_b200"
__path__.insert(0, str(_real))
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
