"""Import alias so that qibo's plugin loader finds this engine under the reference's name.

``qibo.set_backend(backend="qibotn", platform=..., runcard=...)`` imports the module ``qibotn``
and calls ``qibotn.MetaBackend.load(platform=..., runcard=...)``
(``/root/reference/src/qibotn/__init__.py:1-5``, ``backends/__init__.py:15``).  Putting the
repository root on ``sys.path`` (or installing it) therefore makes ``qibotn_b200`` the drop-in;
see INTEGRATION.md."""
from qibotn_b200 import MetaBackend, __version__  # noqa: F401
from qibotn_b200 import backends  # noqa: F401
