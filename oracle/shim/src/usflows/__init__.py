"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED for this package.

CPU restatement (plain PyTorch, dtype-generic so it runs in fp64) of the part
of `aai-institute/USFlows` (import root `src.usflows`, no version pinned by the
reference: `/root/reference/README.md:9-21`, `pyproject.toml:9-10`) that nf4ad
drives.  The USFlows sources are absent from `/root/reference` and cannot be
fetched, so every semantic marked [RECALL] below is a documented choice, not a
verified copy; the reference's own tests hold no golden numbers for it
(SURVEY.md section 8c).  What IS pinned: the reference's in-tree code
(`nf4ad/transforms.py`, `nf4ad/flows.py`) runs unmodified on top of this
package and its outputs are frozen under `tests/golden/`.
"""
