"""Drop-in host-side mirror of the reference's ``core`` package (``/root/reference/core``) for the mesh-token
decode path.  Same module / class / function names so that the reference's ``infer.py`` runs unmodified;
the work underneath is done by the sm_100a CUDA library in ``edgerunner_b200``."""
