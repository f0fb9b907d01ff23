"""Empty stub so `import pycocotools` in the reference's evaler does not fail. Not product code."""
