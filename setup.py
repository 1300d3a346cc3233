"""Packaging of the drop-in: the same two console entry points the reference installs
(reference setup.py:30-35: `svtyper=svtyper.classic:cli`, `svtyper-sso=svtyper.singlesample:cli`).
The native libraries are built in-tree by `python -c 'import __graft_entry__ as g; g.build()'`
(nvcc for libsvgt.so, g++ for libsvgt_pack.so) and shipped as package data."""
from setuptools import setup

setup(
    name="svtyper_b200",
    version="0.7.1+b200.2",
    description="B200-native genotype-likelihood path behind hall-lab/svtyper's sv_genotype / sso_genotype",
    packages=["svtyper_b200"],
    package_data={"svtyper_b200": ["*.so", "csrc/*"]},
    python_requires=">=3.9",
    install_requires=["numpy"],
    entry_points={"console_scripts": ["svtyper=svtyper_b200.classic:cli", "svtyper-sso=svtyper_b200.singlesample:cli",
                                      "svtyper-paste=svtyper_b200.vcf_paste:cli"]},
)
