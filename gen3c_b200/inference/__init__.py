"""Entry points with the reference's module layout: `python -m gen3c_b200.inference.gen3c_single_image ...` mirrors
`cosmos_predict1/diffusion/inference/gen3c_single_image.py`."""
