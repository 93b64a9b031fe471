"""CPU oracle for the vae-tagger encode+tag hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vae_tagger_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the timed CPU baseline.

Pinning status (see DESIGN.md "Oracle"):
  * encoder (``oracle.encoder``): **parity unpinned** by the reference -- the
    arithmetic lives in ``diffusers`` (un-vendored, ``requirements.txt:3``
    ``diffusers>=0.21.0``, config says 0.30.0.dev0) which is not installed and
    not installable here, and the reference ships no golden vectors.  Pinned by
    our own known-answer tests only (parameter count 34 274 208, 106 tensors,
    shapes, key names).
  * tag head / focal loss / wrapper scale-shift (``oracle.head``): **pinned**
    against outputs of the reference's own ``modules.py`` /
    ``improved_losses.py`` / ``diffusers_vae_loader.py`` imported in the build
    container through a stub ``diffusers`` module
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.pt``).
"""
