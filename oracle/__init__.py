"""CPU oracle for the somax QG/SWM time-stepping hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product path (``somax_b200``) never does; it fails
loudly when the CUDA library is missing.

It is a numpy/scipy restatement of the reference algorithm (jejjohnson/somax
v0.0.6, files cited per function as ``ref: path:line`` relative to
``/root/reference``).  The stencil and elliptic arithmetic of the reference
lives in two un-vendored dependencies that are absent from ``/root/reference``
and cannot be imported in this image (no jax/jaxlib wheels, no network):

  * ``finitevolx @ v0.0.39``   (pyproject.toml:32)   - C-grid operators, pv_inversion
  * ``spectraldiffx >= 0.0.10`` (pyproject.toml:33,88) - DST-I Helmholtz solver
  * ``diffrax >= 0.6.0``        (pyproject.toml:29)   - Tsit5 / ConstantStepSize

Their published algorithms are restated here and anchored
on the reference's own call sites and property tests.

PINNED BY THE REFERENCE'S OWN PRINTED OUTPUTS: the reference holds no golden vectors for this path
and cannot be executed here, but it ships EXECUTED tutorial notebooks whose stored cell outputs
are numbers the real implementation printed (Poisson / Helmholtz errors and maxima to 6-7 digits,
the extrema and corner values of a 56 031-step shallow-water spin-up, `eqx.filter_grad` values after
1121 steps).  ``tools/extract_notebook_outputs.py`` copies them into
``tests/golden/reference_notebook_outputs.json``; ``tests/test_oracle_reference_pins.py`` checks
this oracle against them and shows that the other settings of ``oracle.operators.OperatorSpec`` do
not reproduce them.  Still unpinned: ``Difference2D.grad_perp`` (oracle/reparam.py only) and the
mode ordering of ``decompose_vertical_modes`` (mode matrices are inputs on both sides).
``tools/capture_reference.py`` regenerates the ``.npz`` fixtures from the real reference on a machine
that has JAX.
"""
