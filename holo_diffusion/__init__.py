"""Drop-in ``holo_diffusion`` package: the reference's module paths and class names over the B200 kernels.

Put this repository ahead of the reference checkout on ``sys.path`` and the reference's own entry points import the
B200 implementation unchanged:

    experiment.py:73            from holo_diffusion.holo_diffusion_model import HoloDiffusionModel
    generate_samples.py:27-30   from holo_diffusion.utils.checkpoint_utils import load_experiment
                                from holo_diffusion.utils.render_utils.flyaround import render_flyaround

Every class below is a thin facade of the ``holo_diffusion_b200`` implementation.  When the Implicitron config system
(``pytorch3d.implicitron.tools.config``) imports, the facades are ``Configurable`` / ``ReplaceableBase`` classes with the
reference's fields and are entered in its ``registry`` under the reference's names (``@registry.register``, like
/root/reference/holo_diffusion/utils/diffusion_utils.py:41, holo_multipass_ea.py:15,
holo_voxel_grid_implicit_function.py:148, holo_diffusion_model.py:44), so ``registry.get(Unet3DBase, "SimpleUnet3D")``
and the ``*_class_type`` / ``*_args`` keys of the shipped yaml configs resolve to them.  Without pytorch3d the same
names are plain re-exports.  State-dict keys are the reference's (``net_3d._net.*``,
``_implicit_functions.{i}._fn.render_mlp.*``): the facades adopt the implementation's sub-modules under those names.

pytorch3d is not installable in the build image, so the registered path is exercised against ``oracle/pt3d_stub``
(tests/test_shim_cpu.py); against the real library it is untested.
"""
from ._plugin import HAVE_CONFIG  # noqa: F401
