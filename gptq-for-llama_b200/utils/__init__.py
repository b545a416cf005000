"""The two helpers the reference's load_quant imports from its utils package
(llama_inference.py:8; utils/modelutils.py:4-13).  Dataset loaders, export and the GPTQ solver
are outside the hot path and are not provided."""
from .modelutils import DEV, find_layers
