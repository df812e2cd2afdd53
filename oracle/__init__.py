"""CPU oracle for the Ofighters hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  ``ofighters_b200`` must
never import from here.

Contents
--------
disk.py          restatement of ``skimage.draw.disk`` (third-party, un-pinned;
                 call site /root/reference/ofighters/lib/form.py:226)
ref_shim.py      loader that imports the UNMODIFIED reference from
                 /root/reference under sys.modules shims (this container only)
step_py.py       pure-Python restatement of the arena step
step_c.c         batched C restatement of the same rules (fast checker + CPU
                 baseline), built by ``oracle/build.py``
policy_torch.py  torch-CPU fp32 restatement of the bi-head pointer model
                 (/root/reference/ofighters/agents/qlearnIA_V2.py:123-190)
gen_golden.py    runs the real reference and writes tests/golden/*.npz

Parity pinning: the reference holds no golden vectors of its own (SURVEY
section 4), so step_py/step_c are pinned against traces of the reference
itself executed in the build container (tests/golden/, made by gen_golden.py).
The policy oracle is "parity unpinned": Keras/TensorFlow are absent and the
reference ships no weights, so it follows the layer list + Keras defaults only.
"""
