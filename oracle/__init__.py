"""TEST INFRASTRUCTURE — CPU oracle for the rasterizer path. Not imported by the product package."""
