"""CPU oracle for the Mistral Water ocean hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
The product package never imports this.
"""
