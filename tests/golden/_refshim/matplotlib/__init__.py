"""Empty stand-in so `import matplotlib.pyplot as plt` in the reference succeeds
(plotting is out of scope; TEST INFRASTRUCTURE ONLY)."""
