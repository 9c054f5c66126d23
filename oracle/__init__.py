"""CPU oracle for the MoRe4D 4D-STraG hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``more4d_b200/`` may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` — and there only as the checker / the CPU
baseline, never as the thing shipped.  (The development diagnostics under ``tools/`` use it the
same way: as the checker of a probe, or as the CPU timing beside a GPU timing.)

Contents
--------
``ref_import``  in-place importer for the real reference modules under ``/root/reference``
                (authoring container only; the GPU box has no reference tree).  Used by
                ``tests/golden/make_golden.py`` to generate the committed golden vectors and
                by the CPU tests (when the tree is present) to cross-check the restatement.
``dit_oracle``  torch-fp32 restatement of the Wan2.1-DiT forward (WanTransformer4DModel).
``vae_oracle``  torch-fp32 restatement of the causal Wan VAE + trajectory adaptors.
``project_oracle``  numpy-float32 restatement of the z-buffer point projection
                (``render_with_project``); reproduces the real function exactly on the goldens.
``gs_oracle``   numpy restatement of the 3D-Gaussian-splatting forward rasteriser behind
                ``gs_render`` (third-party ``diff_gaussian_rasterization`` is absent from the
                reference tree: PARITY UNPINNED, see its header).
``cpu_baseline``  bounded-sample timing of the DiT oracle for ``bench.py``.

``ref_import`` also carries the stand-ins the goldens need to run the REAL modules: an Euler flow
scheduler and pipeline stubs for ``WanFunControlPipeline.__call__`` (``load_pipeline``), and a
deterministic stub OmniMAE trunk (``load_with_stub_omnimae``) feeding the real ``feature_adapter``
/ resize / repeat front end.  ``dit_oracle.mpm_front_end`` restates that front end.

Parity pinning: the reference ships NO tests, golden vectors or fixtures,
so parity is "unpinned by the reference's own tests".  The restatement is instead pinned
against outputs of the *reference itself*, run in the authoring container through
``ref_import`` and committed under ``tests/golden/`` with the generating script.
"""
