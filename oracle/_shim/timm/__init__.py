"""Three-symbol stand-in for `timm`, used ONLY by oracle/gen_golden.py to import the
reference network in the build container (timm is not installed there and is not part
of /root/reference). Test infrastructure, never on the product path."""
