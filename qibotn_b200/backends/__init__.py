"""Backend loader (reference ``backends/__init__.py:8-55``)."""

from .abstract import QibotnBackend
from .cutensornet import CuTensorNet, raise_error

# "qutensornet" is what the reference's README still calls the platform
# (README.md:85); it is accepted as an alias.  "quimb" keeps the quimb platform's API
# (backends/quimb.py) on this engine; qmatchatea is a separate simulator, out of scope.
PLATFORMS = ("cutensornet", "qutensornet", "quimb")


class MetaBackend:
    """Meta-backend class which takes care of loading the qibotn backends."""

    @staticmethod
    def load(platform: str, runcard: dict = None, **kwargs) -> QibotnBackend:
        if platform == "quimb":
            from .quimb import QuimbBackend

            quimb_backend = kwargs.get("quimb_backend", "numpy")
            contraction_optimizer = kwargs.get("contraction_optimizer", "auto-hq")
            return QuimbBackend(quimb_backend=quimb_backend, contraction_optimizer=contraction_optimizer)
        if platform in PLATFORMS:
            return CuTensorNet(runcard)
        raise_error(NotImplementedError,
                    f"Unsupported platform {platform}, please pick one in {PLATFORMS}")

    def list_available(self) -> dict:
        available = {}
        for platform in PLATFORMS:
            try:
                MetaBackend.load(platform=platform)
                available[platform] = True
            except Exception:
                available[platform] = False
        return available
