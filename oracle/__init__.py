"""CPU oracle for the XMem space-time memory readout (TEST INFRASTRUCTURE ONLY).

Nothing in the product package (``vos_e_sam_b200``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker / the thing
timed on the host cores -- never as the shipped path.
"""
