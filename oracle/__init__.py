"""CPU restatement of the reference's render path (C++ under this directory) and its Python bindings: test infrastructure only."""
