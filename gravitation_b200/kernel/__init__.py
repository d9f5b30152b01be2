# -*- coding: utf-8 -*-
"""kernel package: `_base_` (front-end contract) and the kernels discovered by `lib.load.inventory`"""
