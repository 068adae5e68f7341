"""Drop-in for the reference's top-level `pointnet2` package (only `_ext` is needed):
`import pointnet2._ext as _ext` at /root/reference/modules/third_party/pointnet2/pointnet2_utils.py:22-23.
Use `sceneverse_b200.dropin.install()` to register it under that name."""
