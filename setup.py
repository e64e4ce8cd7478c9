"""pip-installable package (counterpart of the reference's setup.py, which drives CMake for sm_50..sm_75).

  pip install --no-build-isolation .          # builds csrc/ for sm_100a with nvcc, ships the .so inside the package
  python setup.py build_ext --inplace         # in-tree build only (same as python -m graphlearn_for_pytorch_b200.ops.build)

The native core is built by `graphlearn_for_pytorch_b200/ops/build.py` (one target: compute_100a/sm_100a, -lineinfo);
this file only hooks it into setuptools so that wheels contain `_ext/glt_b200_C.so` next to the sources.
"""
import os
import runpy

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = runpy.run_path(os.path.join(HERE, 'graphlearn_for_pytorch_b200', 'ops', 'build.py'))


class BuildNative(Command):
  description = 'compile csrc/ into graphlearn_for_pytorch_b200/_ext/glt_b200_C.so'
  user_options = [('inplace', 'i', 'kept for compatibility: the build is always in-tree'), ('force', 'f', 'rebuild')]

  def initialize_options(self):
    self.inplace, self.force = 1, 0

  def finalize_options(self):
    pass

  def run(self):
    print('native core:', BUILD['build'](verbose=True, force=bool(self.force)))


class BuildPyWithNative(build_py):
  def run(self):
    self.run_command('build_ext')
    super().run()


setup(
  name='graphlearn_for_pytorch_b200',
  version='0.1.0',
  description='Blackwell-native GNN sampling / feature / training-data engine with the API of graphlearn-for-pytorch',
  packages=find_packages(include=['graphlearn_for_pytorch_b200', 'graphlearn_for_pytorch_b200.*']),
  package_data={'graphlearn_for_pytorch_b200': ['_ext/*.so', 'csrc/*.cc', 'csrc/*.h', 'csrc/cpu/*', 'csrc/cuda/*']},
  python_requires='>=3.10',
  install_requires=['torch>=2.8', 'numpy'],
  extras_require={'tables': ['pyarrow'], 'launch': ['pyyaml', 'paramiko']},
  cmdclass={'build_ext': BuildNative, 'build_py': BuildPyWithNative},
  zip_safe=False,
)
