"""zero-inflated-gp_b200: B200-native OnOffSVGP ELBO + gradient path.

The directory name is not a Python identifier; put it on ``sys.path`` and import the host-side mirror of the reference
package, which keeps the reference's import name::

    sys.path.insert(0, "<repo>/zero-inflated-gp_b200")
    from onoffgpf import OnOffSVGP, OnOffLikelihood, RBF

``csrc/`` holds the CUDA kernels and the C ABI (built in-tree into ``libzigp.so``)."""
