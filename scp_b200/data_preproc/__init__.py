"""Mirror of the reference's ``data_preproc`` package for the encode path (data_preprocess, Octree, OctreeCPP)."""
