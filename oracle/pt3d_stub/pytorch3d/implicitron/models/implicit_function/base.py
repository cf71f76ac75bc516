from pytorch3d.implicitron.tools.config import ReplaceableBase


class ImplicitFunctionBase(ReplaceableBase):
    @staticmethod
    def allows_multiple_passes() -> bool:
        return False
