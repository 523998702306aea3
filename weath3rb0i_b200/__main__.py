import sys

from .codec import main

sys.exit(main())
