import sys
from .frontend import main
sys.exit(main())
