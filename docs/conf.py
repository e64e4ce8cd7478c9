# Sphinx configuration of the documentation site (counterpart of the reference's docs/conf.py).
#   pip install sphinx myst-parser furo && sphinx-build -b html docs docs/_build/html
# The pages are the Markdown files of this directory (MyST); the API reference is generated with autodoc from the
# package's docstrings (the native extension is mocked so the site builds on machines without nvcc / a GPU).
import os
import sys

sys.path.insert(0, os.path.abspath('..'))

project = 'graphlearn-for-pytorch-b200'
author = 'graphlearn-for-pytorch-b200 contributors'
release = '0.1.0'

extensions = ['myst_parser', 'sphinx.ext.autodoc', 'sphinx.ext.autosummary', 'sphinx.ext.napoleon', 'sphinx.ext.viewcode']
source_suffix = {'.md': 'markdown', '.rst': 'restructuredtext'}
master_doc = 'index'
exclude_patterns = ['_build']
autodoc_mock_imports = ['graphlearn_for_pytorch_b200._ext', 'pyarrow', 'paramiko']
autodoc_default_options = {'members': True, 'undoc-members': False, 'show-inheritance': True}
autosummary_generate = True
napoleon_google_docstring = True
html_theme = 'furo'
html_title = 'graphlearn-for-pytorch-b200'
myst_enable_extensions = ['colon_fence', 'deflist']
