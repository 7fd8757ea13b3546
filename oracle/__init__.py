"""CPU oracle for the DiffDock score-model hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``diffdock_b200``) never
imports it and has no CPU fallback.

What it is
----------
A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm, function by
function, each citing the reference file:line it follows (paths relative to the upstream
tree gcorso/DiffDock @ b4704d9):

* ``e3nn_lite``     - the subset of e3nn 0.5.x the path uses (Irreps, Wigner-3j, real
                      spherical harmonics, FullyConnectedTensorProduct, FullTensorProduct,
                      BatchNorm).  e3nn is a pinned third-party dependency
                      (requirements.txt:7 ``e3nn==0.5.0`` / environment.yml:23 ``0.5.1``)
                      that is NOT vendored in the reference tree and NOT installed here.
* ``graph_ops``     - torch_scatter 2.1.0 ``scatter`` and torch_cluster 1.6.0
                      ``radius``/``radius_graph`` semantics (requirements.txt:19-21).
* ``layers``, ``tensor_layers``, ``cg_model`` - models/layers.py, models/tensor_layers.py,
                      models/cg_model.py (score mode).
* ``diffusion``, ``sampling`` - utils/diffusion_utils.py, utils/torsion.py:75-90,
                      utils/geometry.py, utils/sampling.py:69-201, utils/so3.py:89-93,
                      utils/torus.py:79-83.
* ``inputs``        - the input side: datasets/process_mols.py:161-202,279-301 (receptor / ligand
                      graph construction incl. torch.cdist's fp32 arithmetic), utils/torsion.py:15-45,
                      utils/inference_utils.py:229-236, datasets/pdbbind.py:215-230; pinned by
                      ``tests/golden/ref_inputs.pt`` (the unmodified reference functions run by
                      ``tests/golden/make_golden_inputs.py``).

Pinning status
--------------
The reference ships no tests, golden vectors or fixtures for this path.
What pins this oracle:

0. Files: e3nn_lite / graph_ops (third-party semantics), layers / tensor_layers / cg_model (score model),
   old_cg_model (confidence model, models/old_cg_model.py), diffusion / sampling (sampler), tables, ref_shims.
1. In-tree reference code run in the authoring container: ``tests/golden/make_golden.py``
   imports the UNMODIFIED reference modules (models/layers.py, models/tensor_layers.py,
   models/cg_model.py, utils/geometry.py, utils/diffusion_utils.py, utils/torsion.py,
   utils/sampling.py, utils/so3.py, utils/torus.py) from /root/reference and records their
   outputs as fixtures under ``tests/golden/``.  The un-installable third-party packages
   are supplied to those imports by ``oracle/ref_shims.py`` (built on ``e3nn_lite`` and
   ``graph_ops``), so the *reference's own wiring* (CGModel.forward, TensorProductConvLayer,
   FasterTensorProduct, the sampler loop, the pose update) is pinned against real
   reference code.
2. ``FasterTensorProduct`` (models/tensor_layers.py:44-122) is self-contained arithmetic in
   the reference tree; the e3nn_lite Clebsch-Gordan/normalisation recipe is pinned to it
   for every l<=1 path including signs.
3. The e3nn / torch_scatter / torch_cluster semantics themselves (l=2 CG signs, the irrep
   sort order of FullTensorProduct) are restated from the published algorithm and checked
   by equivariance and closed-form identities only:  **parity unpinned** for those
   third-party conventions (they permute/sign weight layouts and therefore only matter
   for loading a real checkpoint, which is not in the tree either).
"""
